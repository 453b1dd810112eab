"""torch.profiler view of ONE eager asynchronous UNet step (SDXL 1024^2, 1 GPU): which framework ops launch the torch
element-wise / copy / cat kernels that the ncu launch list shows (shapes and call counts) -- candidates for fusion."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from distrifuser_b200.pipelines import DistriSDXLPipeline  # noqa: E402
from distrifuser_b200.utils import DistriConfig  # noqa: E402

cfg = DistriConfig(height=1024, width=1024, use_cuda_graph=False)
pipe = DistriSDXLPipeline.from_synthetic(cfg, seed=0)
pipe.set_progress_bar_config(disable=True)
unet, si = pipe.pipeline.unet, pipe.static_inputs


def step(counter):
    unet.set_counter(counter)
    unet(si["sample"], si["timestep"], si["encoder_hidden_states"], added_cond_kwargs=si.get("added_cond_kwargs"), return_dict=False)


for c in range(8):
    step(c)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step(8)
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_input_shape=True)
rows = []
for e in ka:
    t = getattr(e, "self_device_time_total", None)
    if t is None:
        t = getattr(e, "self_cuda_time_total", 0)
    if t > 0:
        rows.append((t, e.count, e.key, str(e.input_shapes)[:150]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total self device time {tot / 1e3:.2f} ms")
for t, n, k, shp in rows[:70]:
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}% {n:4d}  {k:34s} {shp}")
