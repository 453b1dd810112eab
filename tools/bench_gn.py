"""Micro-benchmark of df_groupnorm_fwd (fused statistics + normalise + SiLU) at the GroupNorm shapes of an SDXL step: a CUDA
graph of back-to-back launches on ROTATING tensors (footprint > L2), CUDA events over 5 replays.  GB/s = (read + write) / time."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

SHAPES = {  # name: (b, C, h, w)
    "l0_320": (2, 320, 128, 128), "l1_640": (2, 640, 64, 64), "l2_1280": (2, 1280, 32, 32), "up_1920": (2, 1920, 64, 64),
    "up_2560": (2, 2560, 32, 32), "up_960": (2, 960, 128, 128), "l1_320": (2, 320, 64, 64), "l2_640": (2, 640, 32, 32),
    "hires_l0": (1, 320, 480, 480), "hires_n8_l0": (1, 320, 120, 480),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="l0_320,l1_640,l2_1280,up_1920,up_2560,up_960,l1_320,l2_640,hires_n8_l0")
    a = ap.parse_args()
    from distrifuser_b200 import _lib
    L = _lib.lib()
    G = 32
    for name in a.shapes.split(","):
        b, C, h, w = SHAPES[name]
        nbytes = b * C * h * w * 2
        n = max(4, min(48, int(3e8 // (2 * nbytes)) + 1))
        xs = [torch.randn(b, C, h, w, device="cuda", dtype=torch.float16).contiguous(memory_format=torch.channels_last) for _ in range(n)]
        ys = [torch.empty_like(x) for x in xs]
        gw = torch.randn(C, device="cuda", dtype=torch.float16)
        gb = torch.randn(C, device="cuda", dtype=torch.float16)
        scratch = [torch.zeros(L.df_groupnorm_scratch_bytes(b, G, h, w, C), dtype=torch.uint8, device="cuda") for _ in range(n)]

        def run(i):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.df_groupnorm_fwd(_lib.null_comm(), xs[i].data_ptr(), None, 0, ys[i].data_ptr(), gw.data_ptr(), gb.data_ptr(), b, h, w, C,
                                          G, 1e-5, 0, 1, 0, 1, 0, 0, 0, 1, scratch[i].data_ptr(), st), "gn")
        for i in range(n):
            run(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for i in range(n):
                    run(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (5 * n)
        print(f"{name:12s} b={b} C={C} {h}x{w} ({nbytes / 1e6:5.1f} MB): {us:7.1f} us  {2 * nbytes / us / 1e3:7.0f} GB/s   (graph of {n} launches)")


if __name__ == "__main__":
    main()
