"""End-to-end parity of the product path (fp16, sm_100a kernels, peer-memory comm) against golden vectors produced by
the UNMODIFIED reference (fp32 CPU, gloo) on the same seeded tiny-SDXL UNet and inputs.

Tolerance (stated per SURVEY 7 'fp16 statistics'): the reference side is fp32, the product computes in fp16 with
fp32 accumulation; on eps predictions of std ~0.35 we require mean |err| < 4e-3 and max |err| < 4e-2 per step, i.e.
PSNR > 45 dB against the reference output (peak = max |ref|)."""
import os

import pytest
import torch

from oracle import workloads
from mp_product import run_product_unet

pytestmark = pytest.mark.gpu
CASES = {c.name: c for c in workloads.UNET_CASES}


def _check(case, outs, golden_dir):
    gold = torch.load(os.path.join(golden_dir, f"unet_{case.name}.pt"))["outs"]
    for r, per_rank in enumerate(outs):
        for t, (a, b) in enumerate(zip(per_rank, gold)):
            assert a.shape == b.shape                                          # identical latent shapes
            err = (a - b).abs()
            mse = (err ** 2).mean().item()
            psnr = 10 * torch.log10(b.abs().max() ** 2 / max(mse, 1e-20)).item()
            assert err.mean().item() < 4e-3 and err.max().item() < 4e-2 and psnr > 45, \
                f"{case.name} rank{r} step{t}: mean {err.mean():.2e} max {err.max():.2e} psnr {psnr:.1f} dB"


def test_unet_single_gpu(golden_dir):
    case = CASES["sdxl_w1"]
    _check(case, run_product_unet(case), golden_dir)


def test_unet_single_gpu_cuda_graph(golden_dir):
    case = CASES["sdxl_w1"]
    _check(case, run_product_unet(case, use_graph=True), golden_dir)


@pytest.mark.parametrize("name", ["sdxl_w2_nosplit", "sdxl_w4_split", "sdxl_w2_fullsync", "sdxl_w2_stale",
                                  "sdxl_w2_nosync", "sdxl_w2_syncgn", "sdxl_w2_sepgn", "sdxl_w4_nosplit"])
def test_unet_multi_rank(name, golden_dir):
    """world_size > 1: real GPUs when the box has them, otherwise the ranks share cuda:0 through CUDA IPC."""
    case = CASES[name]
    _check(case, run_product_unet(case), golden_dir)


@pytest.mark.multigpu(8)
def test_unet_eight_gpus(golden_dir):
    """cfg2 x patch4 on 8 real GPUs (NVLink peer stores between 8 processes); skipped on smaller boxes."""
    case = CASES["sdxl_w8_split"]
    _check(case, run_product_unet(case), golden_dir)


def test_unet_sd15_multi_rank(golden_dir):
    """SD1.x topology (DistriSDPipeline path): head dims 40/80/160/160, 1x1-conv projections, no added embeddings."""
    case = CASES["sd15_w2_nosplit"]
    _check(case, run_product_unet(case), golden_dir)


def test_unet_sd15_four_patches(golden_dir):
    """SD1.x with n=4 patches (head dims 40 / 80 / 160 over 4 K/V segments; deepest level: one row per rank)."""
    case = CASES["sd15_w4_nosplit"]
    _check(case, run_product_unet(case), golden_dir)


@pytest.mark.multigpu(8)
def test_unet_sd15_eight_gpus(golden_dir):
    """BASELINE configs[4] layout (SD1.x, 8 GPUs = cfg2 x patch4, corrected_async_gn) on 8 real GPUs."""
    case = CASES["sd15_w8_split"]
    _check(case, run_product_unet(case), golden_dir)


def test_full_size_sdxl_unet_step_vs_oracle():
    """BASELINE configs[0]: the FULL SDXL UNet (2.57 B parameters, random init), 512x512, one CFG step, world_size 1 --
    the fp16 sm_100a product path against the fp32 CPU oracle on the same weights and inputs.  Tolerance: the same
    relative bar as the tiny-UNet goldens (mean |err| < 1.2 % and max |err| < 12 % of the output's std; PSNR > 45 dB)."""
    import dataclasses
    from oracle import harness
    case = dataclasses.replace(workloads.UNetCase("sdxl_full_512", family="sdxl", world_size=1, latent=64), steps=1)
    got = run_product_unet(case)[0][0]
    want = harness.run_unet(case, impl="oracle")[0]
    assert got.shape == want.shape == (2, 4, 64, 64)
    err = (got - want).abs()
    std = want.std().item()
    mse = (err ** 2).mean().item()
    psnr = 10 * torch.log10(want.abs().max() ** 2 / max(mse, 1e-20)).item()
    assert torch.isfinite(got).all()
    assert err.mean().item() < 1.2e-2 * std and err.max().item() < 0.12 * std and psnr > 45, \
        f"mean {err.mean():.2e} max {err.max():.2e} std {std:.3f} psnr {psnr:.1f} dB"


def test_unet_multi_rank_cuda_graph(golden_dir):
    case = CASES["sdxl_w2_nosplit"]
    _check(case, run_product_unet(case, use_graph=True), golden_dir)
