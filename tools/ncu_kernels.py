"""One launch of every shipped kernel at the benched (SDXL 1024^2, N=1) shapes inside a cudaProfiler range, for
`ncu --profile-from-start off --set full`.  L2 is flushed before each launch.  tools/ncu_summary.py turns the report into
profiles/r2_ncu_*.txt and profiles/ncu_traffic.json (the `roofline.traffic` source of bench.py)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from distrifuser_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
which = set((sys.argv[1] if len(sys.argv) > 1 else "attn,gn,geglu,ln,bias,publish,linear").split(","))


def profiled(fn):
    fn()                       # warm (module load, attribute set-up) outside the range
    torch.cuda.synchronize()
    flush.zero_()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if "attn" in which:
    shapes = [(2, 4096, 4096, 10, 64), (2, 1024, 1024, 20, 64)]
    if "attn3840" in which:
        shapes.append((1, 3600, 14400, 20, 64))
    for (b, lq, lk, h, d) in shapes:
        Cq = h * d
        q = torch.randn(b, lq, Cq, device="cuda", dtype=torch.float16)
        kv = torch.randn(b, lk, 2 * Cq, device="cuda", dtype=torch.float16)
        out = torch.empty_like(q)
        seg = (C.c_int32 * 8)(*range(8))
        ws_bytes = L.df_attn_workspace_bytes(b, lq, lk, 1, h, d)
        ws = torch.zeros(max(ws_bytes, 1), dtype=torch.uint8, device="cuda")
        profiled(lambda: _lib.check(L.df_attn_fwd(_lib.null_comm(), q.data_ptr(), kv.data_ptr(), out.data_ptr(), None, b, lq, lk, h, d,
                                                  q.stride(1), kv.stride(1), out.stride(1), 1, 0, seg, 0, 0, 0.0,
                                                  ws.data_ptr() if ws_bytes else None, ws_bytes, st), "attn"))
if "gn" in which:
    for (Cc, hh, ww) in [(320, 128, 128), (640, 64, 64), (1280, 32, 32)]:
        b, G = 2, 32
        x = torch.randn(b, Cc, hh, ww, device="cuda", dtype=torch.float16).contiguous(memory_format=torch.channels_last)
        w = torch.randn(Cc, device="cuda", dtype=torch.float16)
        bb = torch.randn(Cc, device="cuda", dtype=torch.float16)
        y = torch.empty_like(x)
        scratch = torch.zeros(L.df_groupnorm_scratch_bytes(b, G, hh, ww, Cc), dtype=torch.uint8, device="cuda")
        profiled(lambda: _lib.check(L.df_groupnorm_fwd(_lib.null_comm(), x.data_ptr(), None, 0, y.data_ptr(), w.data_ptr(), bb.data_ptr(), b,
                                                       hh, ww, Cc, G, 1e-5, 0, 1, 0, 1, 0, 0, 0, 1, scratch.data_ptr(), st), "gn"))
if "geglu" in which:
    for (rows, C4) in [(2 * 4096, 4 * 640), (2 * 1024, 4 * 1280)]:
        yy = torch.randn(rows, 2 * C4, device="cuda", dtype=torch.float16)
        profiled(lambda: ops.geglu(yy))
if "ln" in which:
    for (rows, Cc) in [(2 * 4096, 640), (2 * 1024, 1280)]:
        x = torch.randn(rows, Cc, device="cuda", dtype=torch.float16)
        r = torch.randn(rows, Cc, device="cuda", dtype=torch.float16)
        ln = torch.nn.LayerNorm(Cc).cuda().half()
        profiled(lambda: ops.add_layernorm(x, r, ln))
if "bias" in which:
    for (Cc, hh, ww) in [(320, 128, 128), (1280, 32, 32)]:
        a_ = torch.randn(2, Cc, hh, ww, device="cuda", dtype=torch.float16).contiguous(memory_format=torch.channels_last)
        r_ = torch.randn_like(a_)
        bv = torch.randn(Cc, device="cuda", dtype=torch.float16)
        profiled(lambda: _lib.check(L.df_bias_residual_add(a_.data_ptr(), r_.data_ptr(), bv.data_ptr(), a_.data_ptr(), 2 * hh * ww, Cc, st), "bias"))
if "publish" in which:
    from helpers import LoopbackArena
    n, nbytes = 2, 1 * 2048 * 2 * 1280 * 2                 # K|V of level 1 at n=2: [1, 2048, 2*640] ... 10 MB
    arena = LoopbackArena(n, [nbytes], rank=0)
    src = torch.randn(nbytes // 2, device="cuda", dtype=torch.float16)
    arena.set_clock(pub=1, rd=1)
    profiled(lambda: _lib.check(L.df_slot_publish(arena.comm, src.data_ptr(), 1, nbytes, nbytes, arena.tensor_off[0], arena.slot_bytes[0],
                                                  0, 0b10, 64, st), "publish"))
if "linear" in which:
    # level-2 feed-forward of SDXL at 1024^2 (CFG pair): fused GEGLU projection, and the plain FF2 / to_out shapes
    M, K, D = 2048, 1280, 5120
    x = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(2 * D, K, device="cuda") / K ** 0.5).half()
    bias = torch.randn(2 * D, device="cuda").half()
    blk = ops.geglu_block(M, 2 * D, K)
    wi, bi = ops.geglu_interleave(w, bias, blk)
    profiled(lambda: ops.linear_geglu(x, wi, bi, blk))
    h = torch.randn(M, D, device="cuda").half()
    w2 = (torch.randn(K, D, device="cuda") / D ** 0.5).half()
    b2 = torch.randn(K, device="cuda").half()
    r = torch.randn(M, K, device="cuda").half()
    profiled(lambda: ops.linear(h, w2, b2, r))
    wq = (torch.randn(3 * K, K, device="cuda") / K ** 0.5).half()
    profiled(lambda: ops.linear(x, wq))
print("done")
