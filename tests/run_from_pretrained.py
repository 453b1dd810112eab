"""TEST INFRASTRUCTURE -- drives Distri{SDXL,SD}Pipeline.from_pretrained (reference pipelines.py:20-42,179-200) in a fresh
interpreter whose `diffusers` is tests/fake_diffusers (type-checks the unet against ModelMixin like the real package).
    python tests/run_from_pretrained.py check-bases        # CPU: class hierarchy only
    python tests/run_from_pretrained.py sdxl|sd15           # GPU: from_pretrained -> prepare() -> 3-step __call__"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "fake_diffusers"))

import diffusers  # noqa: E402  (the fake)
import torch  # noqa: E402

from distrifuser_b200.compat.unet_2d_condition import SD15, SDXL  # noqa: E402
from distrifuser_b200.models.base_model import BaseModel  # noqa: E402

TINY_SDXL = dict(SDXL, block_out_channels=(64, 128, 256), transformer_layers_per_block=(1, 1, 2), attention_head_dim=(1, 2, 4),
                 cross_attention_dim=64, addition_time_embed_dim=32, projection_class_embeddings_input_dim=32 * 6 + 48)
TINY_SD15 = dict(SD15, block_out_channels=(80, 160, 320, 320), attention_head_dim=(2, 2, 2, 2), cross_attention_dim=48,
                 norm_num_groups=8)


def main():
    what = sys.argv[1]
    assert issubclass(BaseModel, diffusers.ModelMixin) and issubclass(BaseModel, diffusers.ConfigMixin), BaseModel.__mro__
    if what == "check-bases":
        print("OK bases")
        return
    from distrifuser_b200.pipelines import DistriSDPipeline, DistriSDXLPipeline
    from distrifuser_b200.utils import DistriConfig
    sdxl = what == "sdxl"
    diffusers.UNET_CONFIG.update(TINY_SDXL if sdxl else TINY_SD15)
    cfg = DistriConfig(height=256, width=256, warmup_steps=1)
    cls = DistriSDXLPipeline if sdxl else DistriSDPipeline
    pipe = cls.from_pretrained(cfg, pretrained_model_name_or_path="fake/checkpoint")
    assert [c[0] for c in diffusers.CALLS] == ["unet.from_pretrained", "pipeline.from_pretrained"]
    unet = pipe.pipeline.unet
    assert isinstance(unet, diffusers.ModelMixin) and unet.dtype == torch.float16 and unet.device.type == "cuda"
    lat = pipe(prompt="a photo", num_inference_steps=3, guidance_scale=5.0, generator=torch.Generator().manual_seed(0)).images
    assert lat.shape == (1, 4, 32, 32) and torch.isfinite(lat).all()
    print("OK from_pretrained", what)


if __name__ == "__main__":
    main()
