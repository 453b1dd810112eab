"""Reference: distrifuser/models/base_model.py:8-52 (same attributes and methods).

The reference derives BaseModel from diffusers' (ModelMixin, ConfigMixin): StableDiffusion(XL)Pipeline.from_pretrained(...,
unet=DistriUNetPP) type-checks the component against ModelMixin.  When diffusers is importable the same bases are used
here; without it (this image) a plain nn.Module carries the `dtype` / `device` properties the mixin would provide."""
import torch
from torch import nn

from ..modules.base_module import BaseModule
from ..utils import DistriConfig, PatchParallelismCommManager

try:  # pragma: no cover - depends on the environment
    from diffusers import ConfigMixin, ModelMixin
    _BASES = (ModelMixin, ConfigMixin)
except Exception:
    _BASES = (nn.Module,)


class BaseModel(*_BASES):
    def __init__(self, model: nn.Module, distri_config: DistriConfig):
        super(BaseModel, self).__init__()
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


def _first_param(m: nn.Module):
    for p in m.parameters():
        return p
    for b in m.buffers():
        return b
    return None


if not any(hasattr(b, "dtype") for b in _BASES):                   # ModelMixin.dtype / .device for the nn.Module fallback
    BaseModel.dtype = property(lambda self: getattr(_first_param(self), "dtype", torch.float32))
    BaseModel.device = property(lambda self: getattr(_first_param(self), "device", torch.device("cpu")))
