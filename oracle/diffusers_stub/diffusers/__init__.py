"""TEST INFRASTRUCTURE (oracle) -- a stub of the handful of ``diffusers==0.24.0`` symbols the reference
imports (/root/reference/distrifuser/pipelines.py:2, models/base_model.py:1, modules/pp/attn.py:2-3 ...).
diffusers is pinned by the reference (/root/reference/setup.py:14) but is neither vendored nor installed here.
Never imported by the product package."""
from .models.unet_2d_condition import ConfigMixin, ModelMixin, UNet2DConditionModel  # noqa: F401

__version__ = "0.24.0-oracle-stub"


class StableDiffusionPipeline:  # placeholders: the reference only names them at import time
    pass


class StableDiffusionXLPipeline:
    pass
