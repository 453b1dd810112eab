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


def add_layernorm(x: torch.Tensor, r: torch.Tensor | None, norm: torch.nn.LayerNorm):
    """-> (x + r, LayerNorm(x + r)) in one kernel (r=None: (x, LayerNorm(x)))."""
    assert x.is_cuda and x.dtype == torch.float16
    C = x.shape[-1]
    x = x.contiguous()
    y = torch.empty_like(x)
    if r is not None:
        r = r.contiguous()
        s = torch.empty_like(x)
    else:
        s = x
    _lib.check(_lib.lib().df_add_layernorm(x.data_ptr(), r.data_ptr() if r is not None else None,
                                           s.data_ptr() if r is not None else None, y.data_ptr(), norm.weight.data_ptr(),
                                           norm.bias.data_ptr(), x.numel() // C, C, float(norm.eps),
                                           torch.cuda.current_stream().cuda_stream), "df_add_layernorm")
    return s, y
