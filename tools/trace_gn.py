"""%globaltimer trace of one gn_fused_kernel launch (build with DF_NVCC_FLAGS=-DDF_GN_TRACE): where the time of a GroupNorm goes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from distrifuser_b200 import _lib  # noqa: E402

L = _lib.lib()
names = ["start->stats loop done", "fold + partial + grid barrier", "reduce own sample's partials", "coef (+ peer wait)",
         "publish (last arriver only)", "apply (CTA 0)"]
for (Cc, hh, ww) in [(320, 128, 128), (640, 64, 64), (1280, 32, 32)]:
    b, G = 2, 32
    x = torch.randn(b, Cc, hh, ww, device="cuda", dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cc, device="cuda", dtype=torch.float16)
    bb = torch.randn(Cc, device="cuda", dtype=torch.float16)
    y = torch.empty_like(x)
    scratch = torch.zeros(L.df_groupnorm_scratch_bytes(b, G, hh, ww, Cc), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        _lib.check(L.df_groupnorm_fwd(_lib.null_comm(), x.data_ptr(), None, 0, y.data_ptr(), w.data_ptr(), bb.data_ptr(), b, hh, ww, Cc, G,
                                      1e-5, 0, 1, 0, 1, 0, 0, 0, 1, scratch.data_ptr(), st), "gn")
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    L.df_debug_gn_trace(buf)
    t = [buf[i] for i in range(8)]
    d = [t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5]]    # slots: see GN_TR in gn_fused_kernel
    print(f"C={Cc} {hh}x{ww}: " + "  ".join(f"{n}: {v / 1e3:.1f} us" for n, v in zip(names, d)) + f"   | CTA0 total {(t[6] - t[0]) / 1e3:.1f} us, last CTA end {(t[7] - t[0]) / 1e3:.1f} us")
