"""Micro-benchmark of df_attn_fwd (fmha_fwd_kernel) at the SDXL self-attention shapes; CUDA events, L2 flushed between
iterations.  `--profile` brackets a single launch with cudaProfilerStart/Stop for ncu."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

SHAPES = {  # name: (b, lq, lkv, heads, d)
    "1024_l1": (2, 4096, 4096, 10, 64), "1024_l2": (2, 1024, 1024, 20, 64),
    "3840n4_l2": (1, 3600, 14400, 20, 64), "3840n4_l1": (1, 14400, 57600, 10, 64),
    "2048n2_l1": (1, 8192, 16384, 10, 64), "cross_l2": (2, 1024, 77, 20, 64),
    "1024n4_l2": (1, 256, 1024, 20, 64), "1024n4_l1": (1, 1024, 4096, 10, 64), "1024n2_l2": (1, 512, 1024, 20, 64),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="1024_l1,1024_l2,3840n4_l2,2048n2_l1")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--no-split", action="store_true", help="disable split-KV (A/B)")
    a = ap.parse_args()
    from distrifuser_b200 import _lib
    L = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name in a.shapes.split(","):
        b, lq, lk, h, d = SHAPES[name]
        Cq = h * d
        q = torch.randn(b, lq, Cq, device="cuda", dtype=torch.float16)
        kv = torch.randn(b, lk, 2 * Cq, device="cuda", dtype=torch.float16)
        out = torch.empty_like(q)
        seg = (C.c_int32 * 8)(*range(8))
        ws_bytes = 0 if a.no_split else L.df_attn_workspace_bytes(b, lq, lk, 1, h, d)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream

        def run():
            _lib.check(L.df_attn_fwd(_lib.null_comm(), q.data_ptr(), kv.data_ptr(), out.data_ptr(), None, b, lq, lk, h, d,
                                     q.stride(1), kv.stride(1), out.stride(1), 1, 0, seg, 0, 0, 0.0,
                                     ws.data_ptr() if ws_bytes else None, ws_bytes, st), "df_attn_fwd")
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        if a.profile:
            flush.zero_()
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            run()
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            continue
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        fl = 4.0 * b * lq * lk * Cq
        print(f"{name:10s} b={b} lq={lq} lkv={lk} h={h} d={d}: {ms * 1e3:9.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
