"""Minimal Euler / DDIM steppers with the SD / SDXL scheduler configuration of diffusers==0.24.0 (scaled-linear
betas 0.00085..0.012, 1000 train steps, "leading" spacing, steps_offset 1, epsilon prediction).  Used only when
diffusers is absent; the denoising loop and schedulers are diffusers' business in the reference
 -- they are here so the 50-step latency metric can be measured end to end."""
import numpy as np
import torch


def _alphas_cumprod():
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class EulerDiscreteScheduler:
    order = 1

    def __init__(self):
        self.alphas_cumprod = _alphas_cumprod()
        self.timesteps = None
        self.sigmas = None
        self.init_noise_sigma = 1.0
        self._i = 0

    def set_timesteps(self, num_inference_steps: int, device=None):
        ratio = 1000 // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].astype(np.float64) + 1   # leading, offset 1
        sig = ((1 - self.alphas_cumprod) / self.alphas_cumprod).sqrt().numpy()
        sigmas = np.interp(ts, np.arange(len(sig)), sig)
        self.sigmas = torch.tensor(np.concatenate([sigmas, [0.0]]), dtype=torch.float32)     # host side: no syncs
        self.timesteps = torch.tensor(ts, dtype=torch.float32, device=device)
        self._ts_host = ts
        self.init_noise_sigma = float((self.sigmas.max() ** 2 + 1) ** 0.5)
        self._i = 0

    def scale_model_input(self, sample, t=None):
        s = float(self.sigmas[self._i])
        return sample / ((s * s + 1) ** 0.5)

    def step(self, model_output, t, sample):
        s, s_next = float(self.sigmas[self._i]), float(self.sigmas[self._i + 1])
        prev = sample + model_output.to(sample.dtype) * (s_next - s)    # derivative == eps for epsilon prediction
        self._i += 1
        return (prev,)


class DDIMScheduler:
    order = 1

    def __init__(self):
        self.alphas_cumprod = _alphas_cumprod()
        self.init_noise_sigma = 1.0
        self.timesteps = None
        self._i = 0

    def set_timesteps(self, num_inference_steps: int, device=None):
        ratio = 1000 // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].astype(np.int64) + 1
        self._ts_host = ts
        self._ratio = ratio
        self.timesteps = torch.tensor(ts, dtype=torch.float32, device=device)
        self._i = 0

    def scale_model_input(self, sample, t=None):
        return sample

    def step(self, model_output, t, sample):
        ti = int(self._ts_host[self._i])
        prev_t = ti - self._ratio
        a_t = float(self.alphas_cumprod[ti])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.alphas_cumprod[0])
        eps = model_output.to(sample.dtype)
        x0 = (sample - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        prev = a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps
        self._i += 1
        return (prev,)
