"""Small fused ops on the per-step path that are not reference wrappers (thin shims over the C ABI)."""
import torch

from . import _lib


def geglu(y: torch.Tensor) -> torch.Tensor:
    """hidden * gelu_erf(gate) for y = [..., 2*cols] = [hidden | gate] (diffusers GEGLU.forward after the projection)."""
    assert y.is_cuda and y.dtype == torch.float16 and y.stride(-1) == 1
    cols = y.shape[-1] // 2
    y2 = y.reshape(-1, 2 * cols)
    out = torch.empty((*y.shape[:-1], cols), dtype=y.dtype, device=y.device)
    _lib.check(_lib.lib().df_geglu(y2.data_ptr(), out.data_ptr(), y2.shape[0], cols, y2.stride(0), cols,
                                   torch.cuda.current_stream().cuda_stream), "df_geglu")
    return out
