"""DDIM scheduler with the interface subset of diffusers' DDIMScheduler that the CamAnimate pipelines use
(set_timesteps / timesteps / scale_model_input / step(...).prev_sample / init_noise_sigma / order), for the
noise_scheduler_kwargs of configs/inference/inference_v2.yaml:24-33 (linear betas, zero-terminal-SNR rescale,
trailing spacing, v-prediction, eta=0, no clipping).  Scalars stay on the host, tensors stay where they are."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=False,
                 steps_offset=1, prediction_type="v_prediction", rescale_betas_zero_snr=True, timestep_spacing="trailing", **_):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        if rescale_betas_zero_snr:
            s = torch.cumprod(1.0 - betas, 0).sqrt()
            s0, sT = s[0].clone(), s[-1].clone()
            s = (s - sT) * s0 / (s0 - sT)
            a = s**2
            betas = 1 - torch.cat([a[0:1], a[1:] / a[:-1]])
        if clip_sample:
            raise NotImplementedError("clip_sample")
        self.alphas_cumprod = torch.cumprod(1.0 - betas, 0)
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type, timestep_spacing=timestep_spacing,
                                      steps_offset=steps_offset)
        self.timesteps = None
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps, device=None):
        n, T = num_inference_steps, self.config.num_train_timesteps
        sp = self.config.timestep_spacing
        if sp == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)) - 1
        elif sp == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy() + self.config.steps_offset
        else:
            ts = np.linspace(0, T - 1, n).round()[::-1].copy()
        self.num_inference_steps = n
        self.timesteps = torch.from_numpy(ts.astype(np.int64)).to(device)
        self._host_timesteps = [int(v) for v in ts]

    def coef_table(self):
        """[num_inference_steps][4] fp32: sqrt(a_t), sqrt(1 - a_t), sqrt(a_prev), sqrt(1 - a_prev) per step -- what ``step`` uses,
        laid out for the device-side loop (hv_op_cfg_ddim_step reads row ``*step_index``)."""
        rows = []
        for t in self._host_timesteps:
            prev = t - self.config.num_train_timesteps // self.num_inference_steps
            a_t = float(self.alphas_cumprod[t])
            a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else 1.0
            rows.append([a_t**0.5, (1 - a_t) ** 0.5, a_p**0.5, (1 - a_p) ** 0.5])
        return torch.tensor(rows, dtype=torch.float32)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta=0.0, use_clipped_model_output=False, generator=None, variance_noise=None, return_dict=True):
        """diffusers' ``DDIMScheduler.step`` signature; ``return_dict=False`` returns ``(prev_sample,)`` (pipeline_pose2img.py:351-353 indexes it)."""
        if eta != 0.0:
            raise NotImplementedError("eta != 0")
        t = int(timestep)
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else 1.0
        pt = self.config.prediction_type
        if pt == "v_prediction":
            x0 = (a_t**0.5) * sample - ((1 - a_t) ** 0.5) * model_output
            eps = (a_t**0.5) * model_output + ((1 - a_t) ** 0.5) * sample
        elif pt == "epsilon":
            eps = model_output
            x0 = (sample - ((1 - a_t) ** 0.5) * eps) / (a_t**0.5)
        else:
            raise NotImplementedError(pt)
        prev_sample = (a_p**0.5) * x0 + ((1 - a_p) ** 0.5) * eps
        if not return_dict:
            return (prev_sample.to(sample.dtype),)
        return SimpleNamespace(prev_sample=prev_sample.to(sample.dtype), pred_original_sample=x0)
