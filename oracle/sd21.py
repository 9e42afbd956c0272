"""TEST INFRASTRUCTURE ONLY -- fp32/fp64 torch restatement of the diffusers 0.32.2
modules that `inference_ID-Booth.py:103-138` drives (PARITY UNPINNED, see
oracle/__init__.py).  Pure `torch.nn.functional` ops in NCHW, written from the
published semantics restated in SURVEY.md App. A; every function names the
reference call site whose behaviour it restates.

State dicts use the diffusers key names so a real SD2.1
checkpoint would load.  LoRA adapters are passed as
``{module_path: (down[r,in], up[out,r], scale)}`` and are applied UNMERGED,
``y = x W^T + b + scale * (x A^T) B^T`` (peft `lora.Linear.forward`, injected by
`inference_ID-Booth.py:107`; config `train_ID-Booth.py:672-678`).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Lora = Dict[str, Tuple[Tensor, Tensor, float]]

# ----------------------------------------------------------------------------- configs (App. A.0)
UNET_SD21 = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
    down_attn=(True, True, True, False), up_attn=(False, True, True, True),
    layers_per_block=2, head_dim=64, cross_attention_dim=1024, norm_num_groups=32,
    norm_eps=1e-5, time_embed_in=320,
)
VAE_SD21 = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512),
                layers_per_block=2, norm_num_groups=32, norm_eps=1e-6, scaling_factor=0.18215)


# Attention as an explicit softmax(q k^T / sqrt(d)) v (default: the oracle's arithmetic is spelled out), or through
# `F.scaled_dot_product_attention` -- the call diffusers' AttnProcessor2_0 makes -- when bench.py times this module graph
# on the GPU as the "library" comparator.  Same mathematics either way.
USE_SDPA = False


# ----------------------------------------------------------------------------- primitives
def _linear(sd, name: str, x: Tensor, lora: Optional[Lora] = None) -> Tensor:
    """nn.Linear, optionally wrapped by an unmerged peft LoRA adapter (a8)."""
    w = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    y = F.linear(x, w, b)
    if lora is not None and name in lora:
        down, up, scale = lora[name]
        y = y + scale * F.linear(F.linear(x, down.to(x.dtype)), up.to(x.dtype))
    return y


def _conv(sd, name: str, x: Tensor, stride: int = 1, padding: int = 1) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _gn(sd, name: str, x: Tensor, groups: int, eps: float) -> Tensor:
    return F.group_norm(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps)


def _ln(sd, name: str, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def timestep_sinusoid(t: Tensor, dim: int = 320) -> Tensor:
    """`Timesteps(320, flip_sin_to_cos=True, freq_shift=0)` (a10, App. A.3-T): fp32."""
    half = dim // 2
    k = torch.arange(half, dtype=torch.float32, device=t.device)
    freqs = torch.exp(-math.log(10000.0) * k / half)
    a = t.to(torch.float32)[:, None] * freqs[None, :]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


# ----------------------------------------------------------------------------- blocks (App. A.3)
def resnet_block(sd, p: str, x: Tensor, emb: Optional[Tensor], groups: int, eps: float) -> Tensor:
    """`ResnetBlock2D.forward` (a3)."""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, groups, eps)))
    if emb is not None:
        h = h + _linear(sd, p + ".time_emb_proj", F.silu(emb))[:, :, None, None]
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, groups, eps)))
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def attention(sd, p: str, x: Tensor, ctx: Tensor, heads: int, lora: Optional[Lora]) -> Tensor:
    """`Attention` + `AttnProcessor2_0` (a7): softmax(q k^T / sqrt(d)) v, no mask."""
    q = _linear(sd, p + ".to_q", x, lora)
    k = _linear(sd, p + ".to_k", ctx, lora)
    v = _linear(sd, p + ".to_v", ctx, lora)
    B, T, C = q.shape
    d = C // heads
    q = q.view(B, T, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    if USE_SDPA:
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, C)
    else:
        s = torch.softmax((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(d)), dim=-1)
        o = (s @ v).transpose(1, 2).reshape(B, T, C)
    return _linear(sd, p + ".to_out.0", o, lora)


def basic_transformer_block(sd, p: str, x: Tensor, ctx: Tensor, heads: int, lora) -> Tensor:
    """`BasicTransformerBlock.forward` (a6) with GEGLU feed-forward (a9, exact erf GELU)."""
    n = _ln(sd, p + ".norm1", x)
    x = x + attention(sd, p + ".attn1", n, n, heads, lora)
    x = x + attention(sd, p + ".attn2", _ln(sd, p + ".norm2", x), ctx, heads, lora)
    n = _ln(sd, p + ".norm3", x)
    a, g = _linear(sd, p + ".ff.net.0.proj", n).chunk(2, dim=-1)
    x = x + _linear(sd, p + ".ff.net.2", a * F.gelu(g))
    return x


def transformer_2d(sd, p: str, x: Tensor, ctx: Tensor, heads: int, groups: int, lora) -> Tensor:
    """`Transformer2DModel.forward` (a5), use_linear_projection=True, GN eps 1e-6."""
    B, C, H, W = x.shape
    r = x
    h = _gn(sd, p + ".norm", x, groups, 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = _linear(sd, p + ".proj_in", h)
    h = basic_transformer_block(sd, p + ".transformer_blocks.0", h, ctx, heads, lora)
    h = _linear(sd, p + ".proj_out", h)
    return h.reshape(B, H, W, C).permute(0, 3, 1, 2) + r


# ----------------------------------------------------------------------------- UNet (App. A.2)
def unet_forward(sd, sample: Tensor, timestep, ctx: Tensor, lora: Optional[Lora] = None,
                 cfg: dict = UNET_SD21, taps: Optional[dict] = None) -> Tensor:
    """`UNet2DConditionModel.forward(sample, t, encoder_hidden_states)` (a2;
    signature at `train_ID-Booth.py:1040-1046`).  `taps`, if given, collects
    named intermediates for block-level parity tests."""
    B = sample.shape[0]
    G, eps, hd = cfg["norm_num_groups"], cfg["norm_eps"], cfg["head_dim"]
    ch = cfg["block_out_channels"]
    t = torch.as_tensor(timestep, device=sample.device)
    if t.ndim == 0:
        t = t[None]
    t = t.expand(B)
    temb = timestep_sinusoid(t, cfg["time_embed_in"]).to(sample.dtype)
    emb = _linear(sd, "time_embedding.linear_2", F.silu(_linear(sd, "time_embedding.linear_1", temb)))

    def tap(name, v):
        if taps is not None:
            taps[name] = v.detach().clone()

    h = _conv(sd, "conv_in", sample)
    tap("conv_in", h)
    skips = [h]
    nblk = len(ch)
    for i in range(nblk):
        for j in range(cfg["layers_per_block"]):
            h = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", h, emb, G, eps)
            tap(f"down_blocks.{i}.resnets.{j}", h)
            if cfg["down_attn"][i]:
                h = transformer_2d(sd, f"down_blocks.{i}.attentions.{j}", h, ctx, ch[i] // hd, G, lora)
                tap(f"down_blocks.{i}.attentions.{j}", h)
            skips.append(h)
        if i < nblk - 1:
            h = _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=1)
            skips.append(h)
    h = resnet_block(sd, "mid_block.resnets.0", h, emb, G, eps)
    h = transformer_2d(sd, "mid_block.attentions.0", h, ctx, ch[-1] // hd, G, lora)
    h = resnet_block(sd, "mid_block.resnets.1", h, emb, G, eps)
    tap("mid_block", h)
    rch = tuple(reversed(ch))
    for i in range(nblk):
        for j in range(cfg["layers_per_block"] + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", h, emb, G, eps)
            if cfg["up_attn"][i]:
                h = transformer_2d(sd, f"up_blocks.{i}.attentions.{j}", h, ctx, rch[i] // hd, G, lora)
            tap(f"up_blocks.{i}.{j}", h)
        if i < nblk - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", h)
    h = F.silu(_gn(sd, "conv_norm_out", h, G, eps))
    return _conv(sd, "conv_out", h)


# ----------------------------------------------------------------------------- VAE decoder (App. A.4)
def vae_attention(sd, p: str, x: Tensor, groups: int, eps: float) -> Tensor:
    B, C, H, W = x.shape
    h = _gn(sd, p + ".group_norm", x, groups, eps).view(B, C, H * W).transpose(1, 2)
    q, k, v = (_linear(sd, p + n, h) for n in (".to_q", ".to_k", ".to_v"))
    s = torch.softmax((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(C)), dim=-1)
    o = _linear(sd, p + ".to_out.0", s @ v)
    return o.transpose(1, 2).reshape(B, C, H, W) + x


def vae_decode(sd, z: Tensor, cfg: dict = VAE_SD21, taps: Optional[dict] = None) -> Tensor:
    """`AutoencoderKL.decode(z).sample` (a13; call sites `train_ID-Booth.py:410-412,435-437`);
    z is already divided by scaling_factor."""
    G, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    h = _conv(sd, "post_quant_conv", z, padding=0)
    h = _conv(sd, "decoder.conv_in", h)
    h = resnet_block(sd, "decoder.mid_block.resnets.0", h, None, G, eps)
    h = vae_attention(sd, "decoder.mid_block.attentions.0", h, G, eps)
    h = resnet_block(sd, "decoder.mid_block.resnets.1", h, None, G, eps)
    if taps is not None:
        taps["mid"] = h.clone()
    n = len(cfg["block_out_channels"])
    for i in range(n):
        for j in range(cfg["layers_per_block"] + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, None, G, eps)
        if i < n - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", h)
        if taps is not None:
            taps[f"up{i}"] = h.clone()
    h = F.silu(_gn(sd, "decoder.conv_norm_out", h, G, eps))
    return _conv(sd, "decoder.conv_out", h)


def vae_encode_moments(sd, x: Tensor, cfg: dict = VAE_SD21) -> Tensor:
    """`AutoencoderKL.encode(x)` up to the posterior moments [n, 8, h/8, w/8] = (mean | logvar)
    (`train_ID-Booth.py:1001-1002`: `vae.encode(pixel_values).latent_dist.sample() * scaling_factor`).  diffusers
    Encoder: conv_in -> 4 DownEncoderBlock2D (2 resnets; Downsample2D = F.pad(0,1,0,1) + conv3x3 stride 2 padding 0)
    -> mid (resnet, 1-head attention, resnet) -> GN / SiLU / conv_out -> quant_conv 1x1."""
    G, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    h = _conv(sd, "encoder.conv_in", x)
    n = len(cfg["block_out_channels"])
    for i in range(n):
        for j in range(cfg["layers_per_block"]):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, None, G, eps)
        if i < n - 1:
            h = F.pad(h, (0, 1, 0, 1))
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=0)
    h = resnet_block(sd, "encoder.mid_block.resnets.0", h, None, G, eps)
    h = vae_attention(sd, "encoder.mid_block.attentions.0", h, G, eps)
    h = resnet_block(sd, "encoder.mid_block.resnets.1", h, None, G, eps)
    h = _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", h, G, eps)))
    return _conv(sd, "quant_conv", h, padding=0)


def postprocess_np(image: Tensor):
    """`VaeImageProcessor.postprocess(output_type="np")` (a14; mirrored in-tree at
    `train_ID-Booth.py:413-415`)."""
    return (image * 0.5 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).float().numpy()


# ----------------------------------------------------------------------------- DDPMScheduler (App. A.5)
class DDPMSchedulerRef:
    """`DDPMScheduler` as configured by `DDPMScheduler.from_pretrained(sd21-base,
    subfolder="scheduler")` (`inference_ID-Booth.py:104`): scaled_linear betas,
    leading spacing, steps_offset 1, fixed_small variance, no clipping."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 steps_offset=1, prediction_type="epsilon"):
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                    dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.steps_offset = steps_offset
        self.prediction_type = prediction_type
        self.init_noise_sigma = 1.0
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)
        self.custom = False

    def set_timesteps(self, n: int):
        ratio = self.num_train_timesteps // n
        ts = (torch.arange(n, dtype=torch.float64) * ratio).round().flip(0).to(torch.int64)
        self.timesteps = ts + self.steps_offset
        self.custom = True

    def previous_timestep(self, t: int) -> int:
        if self.custom:
            idx = (self.timesteps == t).nonzero()[0][0].item()
            return -1 if idx == len(self.timesteps) - 1 else int(self.timesteps[idx + 1])
        return t - 1

    def coefficients(self, t: int):
        """(sqrt_acp_t, sqrt_1m_acp_t, c_x0, c_xt, sigma) -- App. C known answers."""
        prev = self.previous_timestep(t)
        acp_t = self.alphas_cumprod[t]
        acp_p = self.alphas_cumprod[prev] if prev >= 0 else torch.tensor(1.0)
        bp_t, bp_p = 1 - acp_t, 1 - acp_p
        a_cur = acp_t / acp_p
        b_cur = 1 - a_cur
        c_x0 = acp_p ** 0.5 * b_cur / bp_t
        c_xt = a_cur ** 0.5 * bp_p / bp_t
        var = torch.clamp(bp_p / bp_t * b_cur, min=1e-20)
        return float(acp_t ** 0.5), float(bp_t ** 0.5), float(c_x0), float(c_xt), float(var ** 0.5)

    def step(self, model_output: Tensor, t: int, sample: Tensor, noise: Optional[Tensor]):
        """Returns (prev_sample, pred_original_sample).  `noise` is the tensor the
        reference would draw with `randn_tensor(..., generator)`; drawn for every
        t > 0 (including the last inference step)."""
        sa, sb, c0, ct, sigma = self.coefficients(int(t))
        if self.prediction_type == "epsilon":
            x0 = (sample - sb * model_output) / sa
        else:  # v_prediction (`train_ID-Booth.py:1057-1058`)
            x0 = sa * sample - sb * model_output
        prev = c0 * x0 + ct * sample
        if int(t) > 0:
            prev = prev + sigma * noise
        return prev, x0

    def add_noise(self, x0: Tensor, noise: Tensor, t: Tensor) -> Tensor:
        acp = self.alphas_cumprod.to(device=x0.device, dtype=x0.dtype)[t.to(x0.device)]
        sa = (acp ** 0.5).view(-1, *([1] * (x0.ndim - 1)))
        sb = ((1 - acp) ** 0.5).view(-1, *([1] * (x0.ndim - 1)))
        return sa * x0 + sb * noise


# ----------------------------------------------------------------------------- pipeline loop (App. A.1)
def denoise_loop(unet_sd, lora, prompt_embeds: Tensor, negative_embeds: Tensor, noise_tape: Tensor,
                 num_steps: int = 30, guidance_scale: float = 5.0, cfg: dict = UNET_SD21,
                 teacher: Optional[Tensor] = None, max_steps: Optional[int] = None,
                 prediction_type: str = "epsilon"):
    """Steps 3-6 of `StableDiffusionPipeline.__call__` (a1; `inference_ID-Booth.py:138`).
    noise_tape[0] = initial latents draw, noise_tape[1+i] = draw of step i.
    Returns the per-step latents [steps+1, n, 4, h, w] (index 0 = initial) and the
    per-step CFG-combined eps.  `teacher`: if given (same shape as the returned
    latents), step i starts from teacher[i] instead of the free-running latent."""
    sch = DDPMSchedulerRef(prediction_type=prediction_type)   # "v_prediction": the 768-v checkpoints (App. A.0)
    sch.set_timesteps(num_steps)
    ctx = torch.cat([negative_embeds, prompt_embeds], dim=0)  # uncond first
    lat = noise_tape[0] * sch.init_noise_sigma
    lats, epss = [lat], []
    for i, t in enumerate(sch.timesteps.tolist()):
        if max_steps is not None and i >= max_steps:
            break
        if teacher is not None:
            lat = teacher[i]
        x2 = torch.cat([lat, lat], dim=0)
        eps = unet_forward(unet_sd, x2, t, ctx, lora, cfg)
        eps_u, eps_c = eps.chunk(2)
        eps = eps_u + guidance_scale * (eps_c - eps_u)
        lat, _ = sch.step(eps, t, lat, noise_tape[1 + i])
        lats.append(lat)
        epss.append(eps)
    return torch.stack(lats), torch.stack(epss)


def merge_lora(sd: dict, lora: Lora) -> dict:
    """W' = W + scale * B A -- used only for the merged == unmerged metamorphic test."""
    out = dict(sd)
    for name, (down, up, scale) in lora.items():
        w = sd[name + ".weight"]
        out[name + ".weight"] = w + scale * (up.to(w.dtype) @ down.to(w.dtype))
    return out
