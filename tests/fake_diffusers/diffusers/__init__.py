"""TEST INFRASTRUCTURE -- a minimal stand-in for the `diffusers` top-level API that DistriSD(XL)Pipeline.from_pretrained
drives (reference: distrifuser/pipelines.py:20-42,179-200), including the component type check diffusers performs on a
user-supplied `unet=` (pipeline_utils.maybe_raise_or_warn: "... is of type ... but should be ModelMixin").  Weights are
random (no checkpoints exist here); the UNet is the product's own diffusers-compatible module tree."""
import torch
from torch import nn

__version__ = "0.24.0-fake"
UNET_CONFIG = {}          # set by the test: constructor overrides of the tiny UNet
CALLS = []


class ConfigMixin:
    config_name = "config.json"


class ModelMixin(nn.Module):
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


class UNet2DConditionModel(ModelMixin, ConfigMixin):
    @classmethod
    def from_pretrained(cls, name, torch_dtype=torch.float32, subfolder=None, **kw):
        from distrifuser_b200.compat.unet_2d_condition import UNet2DConditionModel as Compat
        CALLS.append(("unet.from_pretrained", name, subfolder))
        torch.manual_seed(0)
        return Compat(**UNET_CONFIG).to(torch_dtype)


class _Pipeline:
    sdxl = True

    @classmethod
    def from_pretrained(cls, name, torch_dtype=torch.float32, unet=None, **kw):
        if not isinstance(unet, ModelMixin):
            raise ValueError(f"{unet.__class__.__name__} is of type: {type(unet)}, but should be {ModelMixin}")
        assert unet.dtype == torch_dtype
        from distrifuser_b200.compat.pipeline import SyntheticLatentPipeline
        CALLS.append(("pipeline.from_pretrained", name))
        return SyntheticLatentPipeline(unet, None, sdxl=cls.sdxl, device="cpu", dtype=torch_dtype)


class StableDiffusionXLPipeline(_Pipeline):
    sdxl = True


class StableDiffusionPipeline(_Pipeline):
    sdxl = False
