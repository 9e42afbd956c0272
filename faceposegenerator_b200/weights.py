"""Parameter manifests (diffusers key names) and the deterministic
random-init weight factory used when no SD2.1 snapshot is available offline
(`StableDiffusionPipeline.from_pretrained("stabilityai/stable-diffusion-2-1-base")`,
`/root/reference/inference_ID-Booth.py:103`), plus the LoRA safetensors reader/writer
for the on-disk contract written by `/root/reference/train_ID-Booth.py:696-720`.
"""
from __future__ import annotations

import hashlib
import os
import re
from typing import Dict, Iterable, List, Tuple

import torch

UNET_CONFIG = dict(
    sample_size=64, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    layers_per_block=2, attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024,
    use_linear_projection=True, norm_num_groups=32, norm_eps=1e-5, act_fn="silu",
    flip_sin_to_cos=True, freq_shift=0,
)
VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                  layers_per_block=2, norm_num_groups=32, act_fn="silu", scaling_factor=0.18215, sample_size=512)
SCHEDULER_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                        beta_schedule="scaled_linear", prediction_type="epsilon", clip_sample=False,
                        steps_offset=1, trained_betas=None, variance_type="fixed_small",
                        thresholding=False, timestep_spacing="leading", clip_sample_range=1.0,
                        rescale_betas_zero_snr=False)

Shape = Tuple[int, ...]


def _resnet(p: str, cin: int, cout: int, temb: int | None) -> List[Tuple[str, Shape]]:
    out = [(p + ".norm1.weight", (cin,)), (p + ".norm1.bias", (cin,)),
           (p + ".conv1.weight", (cout, cin, 3, 3)), (p + ".conv1.bias", (cout,))]
    if temb:
        out += [(p + ".time_emb_proj.weight", (cout, temb)), (p + ".time_emb_proj.bias", (cout,))]
    out += [(p + ".norm2.weight", (cout,)), (p + ".norm2.bias", (cout,)),
            (p + ".conv2.weight", (cout, cout, 3, 3)), (p + ".conv2.bias", (cout,))]
    if cin != cout:
        out += [(p + ".conv_shortcut.weight", (cout, cin, 1, 1)), (p + ".conv_shortcut.bias", (cout,))]
    return out


def _transformer(p: str, c: int, cross: int) -> List[Tuple[str, Shape]]:
    t = p + ".transformer_blocks.0"
    out = [(p + ".norm.weight", (c,)), (p + ".norm.bias", (c,)),
           (p + ".proj_in.weight", (c, c)), (p + ".proj_in.bias", (c,))]
    for n in ("norm1", "norm2", "norm3"):
        out += [(f"{t}.{n}.weight", (c,)), (f"{t}.{n}.bias", (c,))]
    for a, kv in (("attn1", c), ("attn2", cross)):
        out += [(f"{t}.{a}.to_q.weight", (c, c)), (f"{t}.{a}.to_k.weight", (c, kv)),
                (f"{t}.{a}.to_v.weight", (c, kv)), (f"{t}.{a}.to_out.0.weight", (c, c)),
                (f"{t}.{a}.to_out.0.bias", (c,))]
    out += [(f"{t}.ff.net.0.proj.weight", (8 * c, c)), (f"{t}.ff.net.0.proj.bias", (8 * c,)),
            (f"{t}.ff.net.2.weight", (c, 4 * c)), (f"{t}.ff.net.2.bias", (c,))]
    out += [(p + ".proj_out.weight", (c, c)), (p + ".proj_out.bias", (c,))]
    return out


def unet_manifest(cfg: dict = UNET_CONFIG) -> List[Tuple[str, Shape]]:
    """Every UNet2DConditionModel parameter (name, shape) in SD2.1-base: 865,910,724 values."""
    ch = tuple(cfg["block_out_channels"])
    temb, cross = ch[0] * 4, cfg["cross_attention_dim"]
    m: List[Tuple[str, Shape]] = [("conv_in.weight", (ch[0], cfg["in_channels"], 3, 3)), ("conv_in.bias", (ch[0],)),
                                  ("time_embedding.linear_1.weight", (temb, ch[0])), ("time_embedding.linear_1.bias", (temb,)),
                                  ("time_embedding.linear_2.weight", (temb, temb)), ("time_embedding.linear_2.bias", (temb,))]
    n = len(ch)
    skip_ch = [ch[0]]
    prev = ch[0]
    for i in range(n):
        attn = cfg["down_block_types"][i].startswith("CrossAttn")
        for j in range(cfg["layers_per_block"]):
            m += _resnet(f"down_blocks.{i}.resnets.{j}", prev, ch[i], temb)
            prev = ch[i]
            if attn:
                m += _transformer(f"down_blocks.{i}.attentions.{j}", ch[i], cross)
            skip_ch.append(ch[i])
        if i < n - 1:
            m += [(f"down_blocks.{i}.downsamplers.0.conv.weight", (ch[i], ch[i], 3, 3)),
                  (f"down_blocks.{i}.downsamplers.0.conv.bias", (ch[i],))]
            skip_ch.append(ch[i])
    m += _resnet("mid_block.resnets.0", prev, prev, temb)
    m += _transformer("mid_block.attentions.0", prev, cross)
    m += _resnet("mid_block.resnets.1", prev, prev, temb)
    rch = tuple(reversed(ch))
    for i in range(n):
        attn = cfg["up_block_types"][i].startswith("CrossAttn")
        for j in range(cfg["layers_per_block"] + 1):
            s = skip_ch.pop()
            m += _resnet(f"up_blocks.{i}.resnets.{j}", prev + s, rch[i], temb)
            prev = rch[i]
            if attn:
                m += _transformer(f"up_blocks.{i}.attentions.{j}", rch[i], cross)
        if i < n - 1:
            m += [(f"up_blocks.{i}.upsamplers.0.conv.weight", (rch[i], rch[i], 3, 3)),
                  (f"up_blocks.{i}.upsamplers.0.conv.bias", (rch[i],))]
    m += [("conv_norm_out.weight", (ch[0],)), ("conv_norm_out.bias", (ch[0],)),
          ("conv_out.weight", (cfg["out_channels"], ch[0], 3, 3)), ("conv_out.bias", (cfg["out_channels"],))]
    return m


def vae_decoder_manifest(cfg: dict = VAE_CONFIG) -> List[Tuple[str, Shape]]:
    """AutoencoderKL decoder + post_quant_conv: 49,490,199 values."""
    ch = tuple(cfg["block_out_channels"])
    lc = cfg["latent_channels"]
    top = ch[-1]
    m: List[Tuple[str, Shape]] = [("post_quant_conv.weight", (lc, lc, 1, 1)), ("post_quant_conv.bias", (lc,)),
                                  ("decoder.conv_in.weight", (top, lc, 3, 3)), ("decoder.conv_in.bias", (top,))]
    m += _resnet("decoder.mid_block.resnets.0", top, top, None)
    a = "decoder.mid_block.attentions.0"
    m += [(a + ".group_norm.weight", (top,)), (a + ".group_norm.bias", (top,))]
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        m += [(f"{a}.{n}.weight", (top, top)), (f"{a}.{n}.bias", (top,))]
    m += _resnet("decoder.mid_block.resnets.1", top, top, None)
    prev = top
    rch = tuple(reversed(ch))
    for i, c in enumerate(rch):
        for j in range(cfg["layers_per_block"] + 1):
            m += _resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev, c, None)
            prev = c
        if i < len(rch) - 1:
            m += [(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (c, c, 3, 3)),
                  (f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (c,))]
    m += [("decoder.conv_norm_out.weight", (prev,)), ("decoder.conv_norm_out.bias", (prev,)),
          ("decoder.conv_out.weight", (cfg["out_channels"], prev, 3, 3)), ("decoder.conv_out.bias", (cfg["out_channels"],))]
    return m


def vae_encoder_manifest(cfg: dict = VAE_CONFIG) -> List[Tuple[str, Shape]]:
    """AutoencoderKL encoder + quant_conv: 34,163,592 + 72 values (decoder + encoder = 83,653,863, the SD VAE)."""
    ch = tuple(cfg["block_out_channels"])
    lc = cfg["latent_channels"]
    m: List[Tuple[str, Shape]] = [("encoder.conv_in.weight", (ch[0], cfg["in_channels"], 3, 3)), ("encoder.conv_in.bias", (ch[0],))]
    prev = ch[0]
    for i, c in enumerate(ch):
        for j in range(cfg["layers_per_block"]):
            m += _resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev, c, None)
            prev = c
        if i < len(ch) - 1:
            m += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (c, c, 3, 3)),
                  (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (c,))]
    top = ch[-1]
    m += _resnet("encoder.mid_block.resnets.0", top, top, None)
    a = "encoder.mid_block.attentions.0"
    m += [(a + ".group_norm.weight", (top,)), (a + ".group_norm.bias", (top,))]
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        m += [(f"{a}.{n}.weight", (top, top)), (f"{a}.{n}.bias", (top,))]
    m += _resnet("encoder.mid_block.resnets.1", top, top, None)
    m += [("encoder.conv_norm_out.weight", (top,)), ("encoder.conv_norm_out.bias", (top,)),
          ("encoder.conv_out.weight", (2 * lc, top, 3, 3)), ("encoder.conv_out.bias", (2 * lc,)),
          ("quant_conv.weight", (2 * lc, 2 * lc, 1, 1)), ("quant_conv.bias", (2 * lc,))]
    return m


LORA_TARGETS = ("to_q", "to_k", "to_v", "to_out.0")  # train_ID-Booth.py:676 (add_k/v_proj match nothing)


def lora_manifest(cfg: dict = UNET_CONFIG, rank: int = 4) -> List[Tuple[str, int, int]]:
    """(module path, in_features, out_features) for the 128 adapted Linears."""
    out = []
    for name, shape in unet_manifest(cfg):
        mm = re.match(r"(.*\.attn[12]\.(to_q|to_k|to_v|to_out\.0))\.weight$", name)
        if mm:
            out.append((mm.group(1), shape[1], shape[0]))
    return out


# ----------------------------------------------------------------------------- deterministic init
def _seed_for(name: str, seed: int) -> int:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return int.from_bytes(h[:8], "little") & 0x7FFF_FFFF_FFFF_FFFF


def _init_tensor(name: str, shape: Shape, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(_seed_for(name, seed))
    if name.endswith(".bias"):
        if ".norm" in name or "group_norm" in name or "conv_norm_out" in name:
            return 0.1 * torch.randn(shape, generator=g)
        return 0.02 * torch.randn(shape, generator=g)
    if len(shape) == 1:  # norm gain
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return torch.randn(shape, generator=g) * (fan_in ** -0.5)


def random_state_dict(manifest: Iterable[Tuple[str, Shape]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Order-independent: each tensor is drawn from a generator keyed by sha256(seed:name).
    Variance-preserving N(0, 1/fan_in) weights keep activations O(1) through 60+ layers."""
    return {name: _init_tensor(name, shape, seed) for name, shape in manifest}


def random_lora(cfg: dict = UNET_CONFIG, rank: int = 4, seed: int = 0, up_std: float = 0.02):
    """A *trained-looking* adapter set: A ~ N(0, (1/r)^2) (peft "gaussian" init,
    train_ID-Booth.py:675) and non-zero B so the LoRA path is exercised."""
    out = {}
    for path, fin, fout in lora_manifest(cfg, rank):
        ga = torch.Generator().manual_seed(_seed_for(path + ".A", seed))
        gb = torch.Generator().manual_seed(_seed_for(path + ".B", seed))
        out[path] = (torch.randn((rank, fin), generator=ga) / rank,
                     torch.randn((fout, rank), generator=gb) * up_std, 1.0)
    return out


# ----------------------------------------------------------------------------- LoRA on-disk format (App. A.6)
LORA_FILE = "pytorch_lora_weights.safetensors"


def save_lora_weights(directory: str, lora: dict, weight_name: str = LORA_FILE) -> str:
    """Writes `unet.<path>.lora.down.weight` / `.lora.up.weight` fp32 tensors, the format
    `LoraLoaderMixin.save_lora_weights` produces at train_ID-Booth.py:716-720,1254-1258."""
    from safetensors.torch import save_file
    os.makedirs(directory, exist_ok=True)
    tensors = {}
    for path, (down, up, scale) in lora.items():
        tensors[f"unet.{path}.lora.down.weight"] = down.detach().float().contiguous()
        tensors[f"unet.{path}.lora.up.weight"] = up.detach().float().contiguous()
        r = down.shape[0]
        if abs(scale - 1.0) > 1e-12:
            tensors[f"unet.{path}.alpha"] = torch.tensor(float(scale * r))
    fn = os.path.join(directory, weight_name)
    save_file(tensors, fn)
    return fn


_LORA_KEY = re.compile(
    r"^(?:unet\.)?(?P<path>.*?)(?:\.processor)?\.(?P<proj>to_q|to_k|to_v|to_out(?:\.0)?)"
    r"(?:_lora\.(?P<legacy>down|up)|\.lora\.(?P<dd>down|up)|\.lora_(?P<peft>A|B)(?:\.[^.]+)?)\.weight$")


def load_lora_state(path_or_dir: str, weight_name: str = LORA_FILE) -> dict:
    """Reads a LoRA checkpoint into {module_path: (down, up, scale)}.  Accepts the
    diffusers spelling (`.lora.down/up`), the peft spelling (`.lora_A/.lora_B[.adapter]`),
    the legacy attn-processor spelling (`.processor.to_q_lora.down`), optional
    `.alpha` scalars (scale = alpha / r; absent => 1.0); `text_encoder.*` tensors are skipped with a warning."""
    from safetensors.torch import load_file
    fn = path_or_dir if os.path.isfile(path_or_dir) else os.path.join(path_or_dir, weight_name)
    if not os.path.isfile(fn):
        raise FileNotFoundError(f"LoRA weights not found: {fn}")
    raw = load_file(fn)
    downs, ups, alphas = {}, {}, {}
    skipped = 0
    for key, t in raw.items():
        if key.startswith("text_encoder."):
            skipped += 1     # `train_text_encoder=True` checkpoints (train_ID-Booth.py:700-707): not applied on this path
            continue
        if key.endswith(".alpha"):
            p = key[:-len(".alpha")]
            p = p[5:] if p.startswith("unet.") else p
            alphas[p.replace(".to_out", ".to_out.0") if p.endswith(".to_out") else p] = float(t)
            continue
        m = _LORA_KEY.match(key)
        if not m:
            raise KeyError(f"unrecognised LoRA key: {key}")
        proj = "to_out.0" if m.group("proj").startswith("to_out") else m.group("proj")
        mod = f"{m.group('path')}.{proj}"
        which = m.group("legacy") or m.group("dd") or {"A": "down", "B": "up"}[m.group("peft")]
        (downs if which == "down" else ups)[mod] = t.float()
    if skipped:
        import warnings
        warnings.warn(f"{fn}: {skipped} text_encoder.* LoRA tensors are NOT applied (the CLIP tower runs without adapters "
                      "on this path); images will differ from a diffusers run that loads them", stacklevel=2)
    if set(downs) != set(ups):
        raise KeyError("LoRA checkpoint has unpaired down/up tensors")
    out = {}
    for mod, d in downs.items():
        r = d.shape[0]
        out[mod] = (d, ups[mod], alphas.get(mod, float(r)) / r)
    return out


# ------------------------------------------------------------------ ArcFace IResNet (config 5)
def iresnet_manifest(arch: str = "r100"):
    """(name, shape) of every state-dict entry of the reference IResNet
    (`/root/reference/ArcFace_files/backbones/iresnet.py:67-162`; iresnet100 = [3, 13, 30, 3]): 925 entries,
    65,156,160 parameters for r100."""
    layers = {"r18": (2, 2, 2, 2), "r34": (3, 4, 6, 3), "r50": (3, 4, 14, 3), "r100": (3, 13, 30, 3)}[arch]
    out = []

    def bn(p, c):
        out.extend([(p + ".weight", (c,)), (p + ".bias", (c,)), (p + ".running_mean", (c,)), (p + ".running_var", (c,)),
                    (p + ".num_batches_tracked", ())])
    out.append(("conv1.weight", (64, 3, 3, 3)))
    bn("bn1", 64)
    out.append(("prelu.weight", (64,)))
    inp = 64
    for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), layers), start=1):
        for b in range(nblk):
            p = f"layer{li}.{b}"
            bn(p + ".bn1", inp)
            out.append((p + ".conv1.weight", (planes, inp, 3, 3)))
            bn(p + ".bn2", planes)
            out.append((p + ".prelu.weight", (planes,)))
            out.append((p + ".conv2.weight", (planes, planes, 3, 3)))
            bn(p + ".bn3", planes)
            if b == 0:
                out.append((p + ".downsample.0.weight", (planes, inp, 1, 1)))
                bn(p + ".downsample.1", planes)
            inp = planes
    bn("bn2", 512)
    out.extend([("fc.weight", (512, 25088)), ("fc.bias", (512,))])
    bn("features", 512)
    return out


def random_iresnet_state_dict(arch: str = "r100", seed: int = 0):
    """Deterministic variance-preserving IResNet weights keyed by parameter name (no checkpoint is available
    offline; the reference loads ArcFace_r100_ms1mv3_backbone.pth, ArcFace_functions.py:29-30)."""
    import hashlib
    sd = {}
    for k, shape in iresnet_manifest(arch):
        g = torch.Generator().manual_seed(int.from_bytes(hashlib.sha256(f"ir{seed}:{k}".encode()).digest()[:8], "little") >> 1)
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("running_var"):
            sd[k] = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif "prelu" in k:
            sd[k] = 0.25 + 0.05 * torch.randn(shape, generator=g)
        elif k.endswith(".bias"):
            sd[k] = 0.05 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            sd[k] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            gain = 0.7 if ("conv2" in k or "downsample" in k) else 1.2
            sd[k] = torch.randn(shape, generator=g) * gain * fan_in ** -0.5
    return sd
