"""`DDPMScheduler` with the diffusers 0.32.2 call surface the reference uses
(`/root/reference/inference_ID-Booth.py:104`, `train_ID-Booth.py:615,1018,1081,1109`):
from_pretrained / config / timesteps / init_noise_sigma / set_timesteps / scale_model_input /
step / add_noise / get_velocity / previous_timestep.

Coefficients are computed once on the host in fp32 exactly as diffusers does (CPU tables) and,
at `set_timesteps`, uploaded as a per-step device table [N, 5] =
(sqrt_acp_t, sqrt_1m_acp_t, c_x0, c_xt, sigma), so `step` is ONE kernel launch
(`idb_cfg_ddpm_step`) with no host synchronisation; the pipeline fuses the CFG combine into
the same launch.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Union

import torch

from .weights import SCHEDULER_CONFIG

f32 = torch.float32


class DDPMSchedulerOutput:
    def __init__(self, prev_sample, pred_original_sample):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample

    def __getitem__(self, i):
        return (self.prev_sample, self.pred_original_sample)[i]


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor: a CPU generator draws on CPU then moves."""
    device = torch.device(device or "cpu")
    gen_device = generator.device if generator is not None else device
    if gen_device.type != device.type and gen_device.type == "cpu":
        return torch.randn(shape, generator=generator, device="cpu", dtype=dtype).to(device)
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)


class DDPMScheduler:
    order = 1

    def __init__(self, **kwargs):
        cfg = dict(SCHEDULER_CONFIG)
        cfg.update({k: v for k, v in kwargs.items() if k in cfg})
        if cfg["beta_schedule"] != "scaled_linear" or cfg["variance_type"] != "fixed_small" or cfg["clip_sample"] \
                or cfg["thresholding"] or cfg["timestep_spacing"] != "leading" or cfg["trained_betas"] is not None:
            raise NotImplementedError("only the SD2.1-base DDPM configuration is implemented")
        self.config = SimpleNamespace(**cfg)
        n = cfg["num_train_timesteps"]
        self.betas = torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5, n, dtype=f32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.custom_timesteps = False
        self.num_inference_steps = None
        self.timesteps = torch.arange(n - 1, -1, -1)
        self._coef_dev = None          # [N, 5] device table for the current timesteps
        self._coef_cache = {}          # (t, device) -> device fp32[5]

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, subfolder=None, **kwargs):
        """Offline: the SD2.1-base scheduler JSON (App. A.0) is built in; a local
        `<dir>/<subfolder>/scheduler_config.json` overrides it when present."""
        import json
        import os
        cfg = {}
        if pretrained_model_name_or_path and os.path.isdir(str(pretrained_model_name_or_path)):
            fn = os.path.join(pretrained_model_name_or_path, subfolder or "", "scheduler_config.json")
            if os.path.isfile(fn):
                with open(fn) as f:
                    cfg = json.load(f)
        cfg.update(kwargs)
        return cls(**cfg)

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = dict(vars(config)) if not isinstance(config, dict) else dict(config)
        cfg.update(kwargs)
        return cls(**cfg)

    def __len__(self):
        return self.config.num_train_timesteps

    # ------------------------------------------------------------------ timesteps
    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device, None] = None):
        n_train = self.config.num_train_timesteps
        if num_inference_steps > n_train:
            raise ValueError("num_inference_steps cannot exceed num_train_timesteps")
        ratio = n_train // num_inference_steps
        ts = (torch.arange(num_inference_steps, dtype=torch.float64) * ratio).round().flip(0).to(torch.int64)
        ts = ts + self.config.steps_offset
        self.num_inference_steps = num_inference_steps
        self.custom_timesteps = False
        self._timesteps_list = ts.tolist()
        self.timesteps = ts.to(device) if device is not None else ts
        self._coef_dev = None
        if device is not None and torch.device(device).type == "cuda":
            rows = [self.coefficients(t) for t in self._timesteps_list]
            self._coef_dev = torch.tensor(rows, dtype=f32).to(device)

    def previous_timestep(self, timestep: int) -> int:
        if self.num_inference_steps:
            lst = self._timesteps_list
            idx = lst.index(int(timestep))
            return -1 if idx == len(lst) - 1 else lst[idx + 1]
        return int(timestep) - 1

    def coefficients(self, timestep: int):
        """fp32 host arithmetic, same operation order as diffusers `DDPMScheduler.step`."""
        t = int(timestep)
        prev = self.previous_timestep(t)
        acp_t = self.alphas_cumprod[t]
        acp_p = self.alphas_cumprod[prev] if prev >= 0 else self.one
        bp_t, bp_p = 1 - acp_t, 1 - acp_p
        a_cur = acp_t / acp_p
        b_cur = 1 - a_cur
        c_x0 = (acp_p ** 0.5 * b_cur) / bp_t
        c_xt = a_cur ** 0.5 * bp_p / bp_t
        var = torch.clamp(bp_p / bp_t * b_cur, min=1e-20)
        sigma = var ** 0.5 if t > 0 else torch.tensor(0.0)
        return [float(acp_t ** 0.5), float(bp_t ** 0.5), float(c_x0), float(c_xt), float(sigma)]

    def coef_row(self, step_index: Optional[int], timestep: int, device) -> torch.Tensor:
        if step_index is not None and self._coef_dev is not None and self._coef_dev.device == torch.device(device):
            return self._coef_dev[step_index]
        key = (int(timestep), str(device), self.num_inference_steps)
        if key not in self._coef_cache:
            self._coef_cache[key] = torch.tensor(self.coefficients(timestep), dtype=f32).to(device)
        return self._coef_cache[key]

    # ------------------------------------------------------------------ API
    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, generator=None, return_dict: bool = True):
        """Ancestral DDPM update (epsilon / v_prediction, fixed_small variance).  Draws
        `randn(model_output.shape, generator, device, dtype)` for t > 0 exactly like diffusers,
        then runs the fused kernel in fp32."""
        from . import ops
        t = int(timestep)
        if not sample.is_cuda:
            raise RuntimeError("DDPMScheduler.step runs on CUDA (sm_100a) only; there is no CPU fallback")
        noise = None
        if t > 0:
            noise = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                                 dtype=model_output.dtype).to(f32).contiguous()
        idx = None
        if self.num_inference_steps and t in self._timesteps_list:
            idx = self._timesteps_list.index(t)
        coef = self.coef_row(idx, t, sample.device)
        x = sample.to(f32).contiguous()
        eps = model_output.to(f32).contiguous()
        x0 = torch.empty_like(x)
        prev = ops.cfg_ddpm_step(eps, x, noise, coef, guidance_scale=1.0, use_cfg=False,
                                 v_prediction=self.config.prediction_type == "v_prediction", x0_out=x0)
        prev, x0 = prev.to(sample.dtype), x0.to(sample.dtype)
        if not return_dict:
            return (prev, x0)
        return DDPMSchedulerOutput(prev, x0)

    def add_noise(self, original_samples, noise, timesteps):
        acp = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = (acp[timesteps] ** 0.5).flatten()
        sb = ((1 - acp[timesteps]) ** 0.5).flatten()
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def get_velocity(self, sample, noise, timesteps):
        acp = self.alphas_cumprod.to(device=sample.device, dtype=sample.dtype)
        timesteps = timesteps.to(sample.device)
        sa = (acp[timesteps] ** 0.5).flatten()
        sb = ((1 - acp[timesteps]) ** 0.5).flatten()
        while sa.dim() < sample.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * noise - sb * sample
