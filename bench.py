#!/usr/bin/env python
"""bench.py -- SDXL 50-step latency (ms/image) of the B200-native patch-parallel UNet path.

    python bench.py --gpus 1 --steps 5 --warmup 3                       # ours, one GPU
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                # the reference's CPU path (oracle port)

A "step" is one full image: 50 denoising steps (CFG pair per step) of the SDXL UNet at --resolution (default
1024x1024 = BASELINE.json configs[1]'s workload; strong scaling: the same image at every N), Euler scheduler,
4 warm-up synchronous steps (DistriConfig defaults), synthetic inputs, random-init weights (no network for
checkpoints).  `value` times the loop with inputs resident in HBM; `e2e` times DistriSDXLPipeline.__call__ with
pinned-host prompt embeddings / latents copied in and the final latents copied out inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT,):
    if p not in sys.path:
        sys.path.insert(0, p)

STEPS_PER_IMAGE = 50
print_json = print
METRIC = "SDXL 50-step latency (ms/image)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--resolution", type=int, default=1024)
    ap.add_argument("--model", default="sdxl", choices=["sdxl", "sd15"])
    ap.add_argument("--mode", default="corrected_async_gn")
    ap.add_argument("--no-split-batch", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exposed-comm", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks line of /opt/skills/guides/B200_PROFILING.md, sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            self.path = tempfile.mktemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm / cpu baseline
FLOPS_STEP = {"sdxl": {512: 3.18e12, 1024: 13.52e12, 2048: 35.91e12 * 2, 3840: 232.46e12 * 2}}   # per CFG pair


def cpu_reference_sample(model: str, resolution: int, reps: int, warm: int):
    """Times the oracle port of the reference's CPU path (fp32, world_size 1: stock UNet forward exactly as
    DistriUNetPP.forward runs it, distri_sdxl_unet_pp.py:118-133) on the host cores: one denoise step (CFG pair) of
    the full-size UNet at 512x512 -- BASELINE.json configs[0], the reference's own CPU-runnable case -- and scales it
    to ms/image of the requested resolution by the per-step FLOP ratio x 50 steps."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_stub"))
    import torch
    from oracle import pp_modules, workloads
    threads = torch.get_num_threads()
    family = "sdxl" if model == "sdxl" else "sd15"
    unet = workloads.make_unet(family, 0)
    cfg = workloads.DuckConfig(1, 0, height=512, width=512)
    wrapped = pp_modules.OracleUNetPP(unet, cfg)
    case = workloads.UNetCase("cpu", family=family, world_size=1, latent=64)
    inp = workloads.unet_inputs(case, 0, workloads.unet_config(family))
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        wrapped(**inp)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    step_s = sum(times) / len(times)
    ratio = FLOPS_STEP["sdxl"].get(resolution, 13.52e12 * (resolution / 1024) ** 2) / FLOPS_STEP["sdxl"][512] if model == "sdxl" else (resolution / 512) ** 2
    ms_image = step_s * ratio * STEPS_PER_IMAGE * 1e3
    sample = (f"{len(times)} timed + {warm} warm-up single denoise steps (CFG pair, fp32) of the full {family} UNet at 512x512 "
              f"on {threads} host threads: {step_s:.2f} s/step; scaled x{ratio:.2f} (FLOP ratio to {resolution}^2) x50 steps")
    return ms_image, threads, sample, step_s


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ms_image, threads, sample, step_s = cpu_reference_sample(a.model, a.resolution, max(1, a.steps), min(a.warmup, 1))
    line = {"metric": METRIC, "value": ms_image, "unit": "ms/image", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_image, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": f"SDXL UNet {a.resolution}x{a.resolution}, 50-step Euler, CFG batch 2, reference CPU path (oracle port, world_size 1)",
                       "inputs_larger_than_l2": True},
            "cpu_baseline": {"value": ms_image, "unit": "ms/image", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": ms_image, "unit": "ms/image", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print_json(json.dumps(line))


# ------------------------------------------------------------------------------------------------ ours
def run_ours(a):
    import torch
    from torch import distributed as dist
    from distrifuser_b200 import _lib
    from distrifuser_b200.compat.schedulers import DDIMScheduler, EulerDiscreteScheduler
    from distrifuser_b200.pipelines import DistriSDPipeline, DistriSDXLPipeline
    from distrifuser_b200.utils import DistriConfig
    _lib.lib()
    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU"
    R = a.resolution
    cfg = DistriConfig(height=R, width=R, mode=a.mode, split_batch=not a.no_split_batch, use_cuda_graph=not a.no_cuda_graph)
    rank, world = cfg.rank, cfg.world_size
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE is {world} (launch with torch.distributed.run)"
    dev = cfg.device
    cls = DistriSDXLPipeline if a.model == "sdxl" else DistriSDPipeline
    pipe = cls.from_synthetic(cfg, seed=0)
    pipe.set_progress_bar_config(disable=True)
    ucfg = pipe.pipeline.unet.config
    B = 2
    g = torch.Generator().manual_seed(1234)                                   # scripts/run_sdxl.py:32
    embeds_h = torch.randn(B, 77, ucfg.cross_attention_dim, generator=g).half().pin_memory()
    pooled_h = None
    if a.model == "sdxl":
        pooled_h = torch.randn(B, ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim, generator=g).half().pin_memory()
    lat_h = torch.randn(1, 4, R // 8, R // 8, generator=g).pin_memory()
    out_h = torch.empty(1, 4, R // 8, R // 8).pin_memory()
    embeds_d, lat_d = embeds_h.to(dev), lat_h.to(dev)
    pooled_d = pooled_h.to(dev) if pooled_h is not None else None

    def image(host: bool):
        kw = dict(num_inference_steps=STEPS_PER_IMAGE, guidance_scale=5.0, output_type="latent")
        if host:
            r = pipe(prompt_embeds=embeds_h, pooled_prompt_embeds=pooled_h, latents=lat_h, **kw)
            out_h.copy_(r.images, non_blocking=True)
        else:
            r = pipe(prompt_embeds=embeds_d, pooled_prompt_embeds=pooled_d, latents=lat_d, **kw)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(host: bool, k: int):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            image(host)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(a.warmup, 3)):
        image(False)
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    n0 = _lib.LAUNCHES["total"]
    total_ms = timed(False, a.steps)
    launches = _lib.LAUNCHES["total"] - n0
    e2e_ms = timed(True, a.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_image = total_ms / a.steps
    ms_image_e2e = e2e_ms / a.steps

    # ---- exposed communication: same kernels with every publication / peer wait removed after warm-up
    #      (mode "no_sync" = compute-only lower bound); (t(mode) - t(no_sync)) / t(mode)
    exposed = None
    if world > 1 and cfg.n_device_per_batch > 1 and not a.no_exposed_comm and a.mode != "no_sync":
        pipe.set_mode("no_sync")
        for _ in range(2):
            image(False)
        nosync_ms = timed(False, a.steps) / a.steps
        pipe.set_mode(a.mode)
        image(False)
        exposed = {"ms_image_no_sync": nosync_ms, "exposed_comm_pct": 100.0 * (ms_image - nosync_ms) / ms_image,
                   "definition": "(t(mode) - t(no_sync)) / t(mode) over the whole 50-step image (5 synchronous + 45 asynchronous steps)"}

    # ---- dominant-kernel roofline (fmha_fwd_kernel, self-attention launches).
    #      (1) one instrumented eager image records the shape of every attention / GroupNorm launch of the model;
    #      (2) every distinct self-attention shape is then timed with CUDA events as a CUDA graph of `count` back-to-back
    #          launches on ROTATING buffers (total footprint > L2), i.e. the kernel's average launch duration at exactly the
    #          step's shapes without the host-launch gaps that eager in-model events pick up for 30-us kernels.
    cfg.use_cuda_graph_saved = cfg.use_cuda_graph
    roof = None
    try:
        cfg.use_cuda_graph = False
        _lib.PROFILE = []
        image(False)
        torch.cuda.synchronize()
        prof, _lib.PROFILE = _lib.PROFILE, None
        cfg.use_cuda_graph = cfg.use_cuda_graph_saved
        gn = [p for p in prof if p["kind"] == "gn"]
        gn_ms = sum(p["start"].elapsed_time(p["end"]) for p in gn)
        gn_b = sum(p["bytes"] for p in gn)
        per_step = {}
        for p in prof:
            if p["kind"] == "self":
                per_step[p["shape"]] = per_step.get(p["shape"], 0) + 1
        per_step = {k: v // STEPS_PER_IMAGE for k, v in per_step.items()}          # launches of that shape per denoise step
        import ctypes as C
        L = _lib.lib()
        seg = (C.c_int32 * 8)(*range(8))
        total_ms_step, total_fl_step, detail = 0.0, 0.0, []
        for (bb, lq_, lkv_, heads_, d_), count in per_step.items():
            Cq = heads_ * d_
            qs = [torch.randn(bb, lq_, Cq, device=dev, dtype=torch.float16) for _ in range(count)]
            kvs = [torch.randn(bb, lkv_, 2 * Cq, device=dev, dtype=torch.float16) for _ in range(count)]
            outs = [torch.empty_like(q_) for q_ in qs]
            side = torch.cuda.Stream(device=dev)

            def launch_all():
                st = torch.cuda.current_stream().cuda_stream
                for q_, kv_, o_ in zip(qs, kvs, outs):
                    _lib.check(L.df_attn_fwd(_lib.null_comm(), q_.data_ptr(), kv_.data_ptr(), o_.data_ptr(), None, bb, lq_, lkv_,
                                             heads_, d_, q_.stride(1), kv_.stride(1), o_.stride(1), 1, 0, seg, 0, 0, 0.0, None, 0,
                                             st), "df_attn_fwd")
            with torch.cuda.stream(side):
                launch_all()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                launch_all()
            for _ in range(2):
                g.replay()
            reps = 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms_launch = e0.elapsed_time(e1) / reps / count
            fl = 4.0 * bb * lq_ * lkv_ * Cq
            total_ms_step += ms_launch * count
            total_fl_step += fl * count
            detail.append({"shape": {"b": bb, "lq": lq_, "lkv": lkv_, "heads": heads_, "d": d_}, "launches_per_step": count,
                           "avg_launch_ms": ms_launch, "tflops": fl / ms_launch / 1e9,
                           "footprint_mb": count * (2 * bb * lq_ * Cq + bb * lkv_ * 2 * Cq) * 2 / 1e6})
            del qs, kvs, outs, g
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1590.0 * 1395.4 / 1700.9)
        ach = total_fl_step / (total_ms_step * 1e-3) / 1e12 if total_ms_step > 0 else 0.0
        n_launch = sum(per_step.values())
        roof = {"kernel": "fmha_fwd_kernel (self-attention launches of one denoise step)", "bound": "tensor", "achieved": ach,
                "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                # dram__bytes_read+write of ONE level-1 launch (b=2, Lq=Lkv=4096, 10 heads) from the ncu --set full capture in
                # profiles/r1_fmha_lvl1.txt; its algorithmic bytes are 2*b*(2*Lq*C + 2*Lkv*C) = 41.9 MB (the O write stays in L2)
                "traffic": 33.8e6 if (a.model == "sdxl" and R == 1024 and world == 1) else None,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback",
                "launches_per_step": n_launch, "avg_launch_ms": total_ms_step / max(n_launch, 1),
                "ms_per_step": total_ms_step, "share_of_step": total_ms_step / (ms_image / STEPS_PER_IMAGE),
                "method": "CUDA events around a CUDA graph of the step's launches of each shape, rotating buffers (> L2)",
                "shapes": detail,
                "groupnorm": {"bound": "hbm", "achieved": gn_b / (gn_ms * 1e-3) / 1e9 if gn_ms > 0 else 0.0,
                              "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s", "launches": len(gn), "ms_per_image": gn_ms,
                              "method": "eager in-model CUDA events (includes launch gaps for the small tensors)"}}
        roof["groupnorm"]["frac"] = roof["groupnorm"]["achieved"] / roof["groupnorm"]["peak"]
    finally:
        cfg.use_cuda_graph = cfg.use_cuda_graph_saved
        _lib.PROFILE = None

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:        # reported at N=1 only (rank 0's host cores)
        v, threads, sample, _ = cpu_reference_sample(a.model, R, 1, 1)
        cpu = {"value": v, "unit": "ms/image", "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        n, b = cfg.n_device_per_batch, (1 if (cfg.do_classifier_free_guidance and cfg.split_batch and world > 1) else 2)
        h2d = embeds_h.numel() * 2 + (pooled_h.numel() * 2 if pooled_h is not None else 0) + lat_h.numel() * 4
        line = {"metric": METRIC, "value": ms_image, "unit": "ms/image", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms_image, "ms_per_denoise_step": ms_image / STEPS_PER_IMAGE, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": f"{a.model.upper()} UNet {R}x{R}, 50-step Euler, CFG batch 2, random-init weights",
                           "parallelism": f"cfg{2 if b == 1 else 1} x patch{n}", "mode": cfg.mode, "warmup_steps": cfg.warmup_steps,
                           "cuda_graph": cfg.use_cuda_graph, "l2": "working set (5.1 GB of weights per step) exceeds the 126 MB L2; no explicit flush"},
                "roofline": roof, "cpu_baseline": cpu, "exposed_comm": exposed,
                "e2e": {"value": ms_image_e2e, "unit": "ms/image", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_h.numel() * 4},
                "gpu_launches": launches, "clocks": clocks}
        print_json(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    # exactly ONE line on stdout: libraries (e.g. "NCCL version ..." at communicator creation) write there too, so the real
    # stdout is parked and fd 1 points at stderr until the JSON line is emitted
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real_stdout, "w")

    def emit(line: str):
        out.write(line + "\n")
        out.flush()

    global print_json
    print_json = emit
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
