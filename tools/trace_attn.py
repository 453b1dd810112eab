"""Cycle-level event trace of one CTA of fmha_fwd_kernel (build with DF_NVCC_FLAGS=-DDF_TRACE).  Prints, per K/V tile,
the clock64 deltas between the events of softmax warp 0 and of the MMA thread."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from distrifuser_b200 import _lib  # noqa: E402

L = _lib.lib()
b, lq, lk, h, d = (int(x) for x in (sys.argv[1:6] or (1, 3600, 14400, 20, 64)))
q = torch.randn(b, lq, h * d, device="cuda", dtype=torch.float16)
kv = torch.randn(b, lk, 2 * h * d, device="cuda", dtype=torch.float16)
out = torch.empty_like(q)
seg = (C.c_int32 * 8)(*range(8))
for _ in range(2):
    _lib.check(L.df_attn_fwd(_lib.null_comm(), q.data_ptr(), kv.data_ptr(), out.data_ptr(), None, b, lq, lk, h, d, q.stride(1),
                             kv.stride(1), out.stride(1), 1, 0, seg, 0, 0, 0.0, None, 0, torch.cuda.current_stream().cuda_stream), "attn")
torch.cuda.synchronize()
buf = (C.c_longlong * (64 * 16))()
L.df_debug_read_trace(buf)
ev = [[buf[t * 16 + s] for s in range(16)] for t in range(64)]
names = ["s_full", "ldtm+free", "max+xchg", "exp", "pv_wait", "sttm+arrive"]
print("tile | softmax warp0: " + "  ".join(f"{n:>11s}" for n in names[1:]) + " | next s_full wait | tile period || MMA: s_free->qk_issued  p_full_at  pv_issued")
for t in range(2, 40):
    e, n = ev[t], ev[t + 1]
    if not e[0] or not n[0]:
        break
    d_ = [e[i + 1] - e[i] for i in range(5)]
    print(f"{t:4d} | " + "  ".join(f"{x:11d}" for x in d_) + f" | {n[0] - e[5]:16d} | {n[0] - e[0]:11d} || {e[9] - e[8]:10d} {e[10] - e[0]:10d} {e[11] - e[10]:10d}   s_free@{e[8] - e[0]}")
