"""Multi-rank driver of the PRODUCT path for the parity tests and multi-GPU runs (test infrastructure).

One process per rank; gloo is only the rendezvous plane (IPC-handle exchange + barriers), every activation moves
through the CUDA peer-memory kernels.  With fewer GPUs than ranks the ranks share cuda:0
(DISTRIFUSER_B200_SHARE_GPU=1): CUDA IPC works between processes on one device, so a 1-GPU box still exercises
the multi-rank slots / flags / epochs (slowly: spin-waits are time-sliced)."""
from __future__ import annotations

import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle", "diffusers_stub"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _worker(rank, case, port, outdir, use_graph):
    from torch import distributed as dist
    world = case.world_size
    if world > 1:
        if torch.cuda.device_count() < world:
            os.environ["DISTRIFUSER_B200_SHARE_GPU"] = "1"
        os.environ["LOCAL_RANK"] = str(rank)
        dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    from oracle import workloads as W
    from distrifuser_b200.compat.unet_2d_condition import UNet2DConditionModel
    from distrifuser_b200.pipelines import DistriSDPipeline, DistriSDXLPipeline
    from distrifuser_b200.utils import DistriConfig
    cfg = DistriConfig(height=8 * case.latent, width=8 * case.latent, do_classifier_free_guidance=case.cfg,
                       split_batch=case.split_batch, warmup_steps=case.warmup_steps, mode=case.mode,
                       use_cuda_graph=use_graph)
    ucfg = W.unet_config(case.family)
    ref_unet = W.make_unet(case.family, case.weight_seed)                  # same seeded weights as the golden run
    unet = UNet2DConditionModel(**ucfg)
    missing = unet.load_state_dict(ref_unet.state_dict(), strict=True)
    del ref_unet
    cls = DistriSDXLPipeline if ucfg.get("addition_embed_type") == "text_time" else DistriSDPipeline
    pipe = cls.from_synthetic(cfg, unet=unet)
    model = pipe.pipeline.unet
    outs = []
    with torch.no_grad():
        model.set_counter(0)                                               # pipelines.py:57
        for t in range(case.steps):
            inp = W.unet_inputs(case, t, ucfg)
            dev = lambda x: x.to(cfg.device, torch.float16) if x.is_floating_point() else x.to(cfg.device)
            kw = dict(sample=dev(inp["sample"]), timestep=inp["timestep"].to(cfg.device).float(),
                      encoder_hidden_states=dev(inp["encoder_hidden_states"]))
            if inp["added_cond_kwargs"] is not None:
                kw["added_cond_kwargs"] = {k: dev(v) for k, v in inp["added_cond_kwargs"].items()}
            outs.append(model(**kw, return_dict=False)[0].float().cpu().clone())
    torch.cuda.synchronize()
    torch.save(outs, os.path.join(outdir, f"rank{rank}.pt"))
    if world > 1:
        dist.barrier()
        if pipe.comm_manager is not None:
            pipe.comm_manager.close()
        dist.destroy_process_group()


def run_product_unet(case, use_graph=False):
    from oracle.harness import free_port
    from torch import multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        if case.world_size == 1:
            _worker(0, case, 0, d, use_graph)
        else:
            mp.spawn(_worker, args=(case, free_port(), d, use_graph), nprocs=case.world_size, join=True)
        return [torch.load(os.path.join(d, f"rank{r}.pt")) for r in range(case.world_size)]


def _traj_worker(rank, case, port, outdir, num_steps, guidance, use_graph):
    from torch import distributed as dist
    world = case.world_size
    if world > 1:
        if torch.cuda.device_count() < world:
            os.environ["DISTRIFUSER_B200_SHARE_GPU"] = "1"
        os.environ["LOCAL_RANK"] = str(rank)
        dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    from oracle import workloads as W
    from distrifuser_b200.compat.unet_2d_condition import UNet2DConditionModel
    from distrifuser_b200.pipelines import DistriSDPipeline, DistriSDXLPipeline
    from distrifuser_b200.utils import DistriConfig
    cfg = DistriConfig(height=8 * case.latent, width=8 * case.latent, do_classifier_free_guidance=case.cfg,
                       split_batch=case.split_batch, warmup_steps=case.warmup_steps, mode=case.mode, use_cuda_graph=use_graph)
    ucfg = W.unet_config(case.family)
    unet = UNet2DConditionModel(**ucfg)
    unet.load_state_dict(W.make_unet(case.family, case.weight_seed).state_dict(), strict=True)
    cls = DistriSDXLPipeline if ucfg.get("addition_embed_type") == "text_time" else DistriSDPipeline
    pipe = cls.from_synthetic(cfg, unet=unet)
    g = torch.Generator().manual_seed(case.input_seed)
    lat = pipe(prompt="a photo", num_inference_steps=num_steps, guidance_scale=guidance, generator=g).images     # public API
    # a second image with the same seed must reproduce the first bit for bit: nothing (epoch banks, text-KV cache, stale
    # activations, graph state) may leak from one image into the next (pipelines.py:57 resets the counters)
    g2 = torch.Generator().manual_seed(case.input_seed)
    lat2 = pipe(prompt="a photo", num_inference_steps=num_steps, guidance_scale=guidance, generator=g2).images
    torch.cuda.synchronize()
    assert torch.equal(lat, lat2), "second image with the same seed differs from the first"
    torch.save(lat.float().cpu(), os.path.join(outdir, f"rank{rank}.pt"))
    if world > 1:
        dist.barrier()
        if pipe.comm_manager is not None:
            pipe.comm_manager.close()
        dist.destroy_process_group()


def run_product_trajectory(case, num_steps=8, guidance=5.0, use_graph=True):
    from oracle.harness import free_port
    from torch import multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        if case.world_size == 1:
            _traj_worker(0, case, 0, d, num_steps, guidance, use_graph)
        else:
            mp.spawn(_traj_worker, args=(case, free_port(), d, num_steps, guidance, use_graph), nprocs=case.world_size, join=True)
        return [torch.load(os.path.join(d, f"rank{r}.pt")) for r in range(case.world_size)]
