"""Reference: distrifuser/models/base_model.py:8-52 (same attributes and methods; ModelMixin/ConfigMixin are
diffusers types and only matter for from_pretrained, so a plain nn.Module is used when diffusers is absent)."""
from torch import nn

from ..modules.base_module import BaseModule
from ..utils import DistriConfig, PatchParallelismCommManager


class BaseModel(nn.Module):
    def __init__(self, model: nn.Module, distri_config: DistriConfig):
        super().__init__()
        self.model = model
        self.distri_config = distri_config
        self.comm_manager = None
        self.buffer_list = None
        self.output_buffer = None
        self.counter = 0
        # for cuda graph
        self.static_inputs = None
        self.static_outputs = None
        self.cuda_graphs = None
        self.graph_launches = None       # number of this package's kernels inside each captured graph

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def set_counter(self, counter: int = 0):                        # base_model.py:27-31
        self.counter = counter
        for module in self.model.modules():
            if isinstance(module, BaseModule):
                module.set_counter(counter)

    def set_comm_manager(self, comm_manager: PatchParallelismCommManager):   # base_model.py:33-37
        self.comm_manager = comm_manager
        for module in self.model.modules():
            if isinstance(module, BaseModule):
                module.set_comm_manager(comm_manager)

    def setup_cuda_graph(self, static_outputs, cuda_graphs, graph_launches=None):   # base_model.py:39-41
        self.static_outputs = static_outputs
        self.cuda_graphs = cuda_graphs
        self.graph_launches = graph_launches

    @property
    def config(self):
        return self.model.config

    def synchronize(self):                                         # base_model.py:47-52
        if self.comm_manager is not None:
            self.comm_manager.join()
