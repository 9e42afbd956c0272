"""ArcFace IResNet (iresnet100 = [3, 13, 30, 3]) in eval mode on the hand-written sm_100a kernels -- the
identity-embedding half of config 5.

Reference: `/root/reference/ArcFace_files/backbones/iresnet.py:29-64` (IBasicBlock), `:67-162` (IResNet),
`:187-189` (iresnet100); loaded frozen by `ArcFace_files/ArcFace_functions.py:27-37`; fed by
`train_ID-Booth.py:433-455` (decode -> crop -> bilinear 112x112 -> (x/255 - 0.5)/0.5) and compared by cosine /
triplet loss at `train_ID-Booth.py:1093-1133`.

Mapping onto `idb_gemm_conv` (NHWC activations, fp32 residual stream, fp16 conv operands like the reference's fp16
autocast, fp32 accumulation):
  * every BatchNorm that FOLLOWS a conv (bn2, bn3, downsample.1, the stem's bn1, `features` after fc) is folded
    into that conv's weights / bias at load (exact in eval mode);
  * a BatchNorm that PRECEDES a conv (IBasicBlock.bn1, IResNet.bn2) cannot be folded (zero padding happens after
    it), so it is the operand producer: `idb_channel_affine` reads the fp32 stream and writes the fp16 operand;
  * PReLU rides in the conv1 epilogue (`prelu` slopes), the block's shortcut add in the conv2 epilogue (`residual`);
  * the stride-2 3x3 conv uses the kernel's 5-D TMA view, the 1x1 stride-2 `downsample` conv is a stride-2
    `idb_channel_affine` (plain sampling) followed by a 1x1 GEMM;
  * fc consumes the NHWC flatten, so its weight columns are permuted from (c, h, w) to (h, w, c) order at load.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops

f16, f32 = torch.float16, torch.float32   # the reference runs the backbone under fp16 autocast (iresnet.py:149)
LAYERS = {"r18": (2, 2, 2, 2), "r34": (3, 4, 6, 3), "r50": (3, 4, 14, 3), "r100": (3, 13, 30, 3)}
STEM_CPAD = 64   # the 3 input channels are zero-padded to one 64-channel K block
EPS = 1e-5


def _bn_affine(sd, p):
    """eval-mode BatchNorm as y = x * scale + shift (fp64 fold)."""
    g, b = sd[p + ".weight"].double(), sd[p + ".bias"].double()
    m, v = sd[p + ".running_mean"].double(), sd[p + ".running_var"].double()
    scale = g / torch.sqrt(v + EPS)
    return scale, b - m * scale


def _fold_conv(w, scale, shift):
    """conv (no bias) followed by BN(scale, shift) -> ([Cout, kh*kw*Cin] tap-major / channel-minor fp64, bias)."""
    w = w.double() * scale.view(-1, 1, 1, 1)
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), shift


class _Block:
    __slots__ = ("s1", "h1", "w1", "b1", "slope", "w2", "b2", "stride", "wd", "bd")


class IResNet:
    def __init__(self, state_dict: Dict[str, torch.Tensor], arch: str = "r100", device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("IResNet runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.arch = arch
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        dev = self.device

        def d32(t):
            return t.to(device=dev, dtype=f32).contiguous()

        def d16(t):
            return t.to(device=dev, dtype=f16).contiguous()

        # stem: conv3x3(3 -> 64) + bn1 folded, PReLU in the epilogue; input channels padded to STEM_CPAD
        sc, sh = _bn_affine(sd, "bn1")
        w = sd["conv1.weight"].double() * sc.view(-1, 1, 1, 1)                 # [64, 3, 3, 3]
        wp = torch.zeros(w.shape[0], 3, 3, STEM_CPAD, dtype=torch.float64)
        wp[..., :3] = w.permute(0, 2, 3, 1)
        self.stem_w, self.stem_b, self.stem_slope = d16(wp.reshape(w.shape[0], -1)), d32(sh), d32(sd["prelu.weight"])
        self.blocks = []
        for li, nblk in enumerate(LAYERS[arch], start=1):
            for bi in range(nblk):
                p = f"layer{li}.{bi}"
                k = _Block()
                s1, h1 = _bn_affine(sd, p + ".bn1")
                k.s1, k.h1 = d32(s1), d32(h1)
                w1, b1 = _fold_conv(sd[p + ".conv1.weight"], *_bn_affine(sd, p + ".bn2"))
                k.w1, k.b1, k.slope = d16(w1), d32(b1), d32(sd[p + ".prelu.weight"])
                w2, b2 = _fold_conv(sd[p + ".conv2.weight"], *_bn_affine(sd, p + ".bn3"))
                k.w2, k.b2 = d16(w2), d32(b2)
                k.stride = 2 if bi == 0 else 1
                if (p + ".downsample.0.weight") in sd:
                    wd, bd = _fold_conv(sd[p + ".downsample.0.weight"], *_bn_affine(sd, p + ".downsample.1"))
                    k.wd, k.bd = d16(wd), d32(bd)
                else:
                    k.wd = k.bd = None
                self.blocks.append(k)
        s2, h2 = _bn_affine(sd, "bn2")
        self.s2, self.h2 = d32(s2), d32(h2)
        # fc + `features` BatchNorm1d folded; columns (c, h, w) -> (h, w, c)
        fs, fh = _bn_affine(sd, "features")
        wf = sd["fc.weight"].double()                                           # [512, C*7*7]
        cl = wf.shape[1] // 49
        wf = wf.view(-1, cl, 7, 7).permute(0, 2, 3, 1).reshape(wf.shape[0], -1) * fs.view(-1, 1)
        self.fc_w, self.fc_b = d16(wf), d32(sd["fc.bias"].double() * fs + fh)
        self._ws = torch.empty((32 << 20) // 4, dtype=f32, device=dev)

    # ------------------------------------------------------------------ forward
    def forward_nhwc(self, x_bf16: torch.Tensor) -> torch.Tensor:
        """x: fp16 [n, 112, 112, STEM_CPAD] (channels 0-2 = RGB in [-1, 1], rest zero) -> fp32 [n, 512]."""
        g = lambda a, w, **kw: ops.gemm_conv(a, w, k_splits=0, workspace=self._ws, half=True, **kw)
        n, H, W, _ = x_bf16.shape
        x, _ = g(x_bf16, self.stem_w, mode=ops.A_3X3, bias=self.stem_b, prelu=self.stem_slope, want_f32=True)
        x = x.view(n, H, W, -1)
        for k in self.blocks:
            n_, H, W, Cin = x.shape
            a = ops.channel_affine(x, k.s1, k.h1, half=True)                                                  # bn1
            _, t = g(a, k.w1, mode=ops.A_3X3, bias=k.b1, prelu=k.slope, want_bf16=True)            # conv1 + bn2 + PReLU
            t = t.view(n_, H, W, -1)
            if k.wd is not None:                                                                   # downsample: conv1x1(stride) + BN
                xs = ops.channel_affine(x, stride=k.stride, half=True)
                sc, _ = g(xs, k.wd, bias=k.bd, want_f32=True)
            else:
                sc = x.view(-1, Cin)
            mode = ops.A_3X3_S2 if k.stride == 2 else ops.A_3X3
            o, _ = g(t, k.w2, mode=mode, bias=k.b2, residual=sc, want_f32=True)                    # conv2 + bn3 + shortcut
            x = o.view(n_, H // k.stride, W // k.stride, -1)
        a = ops.channel_affine(x, self.s2, self.h2, half=True)                                                # bn2
        y, _ = g(a.view(x.shape[0], -1), self.fc_w, bias=self.fc_b, want_f32=True)                 # fc + features BN
        return y

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [n, 3, 112, 112] in [-1, 1] (the reference module's input) -> [n, 512] embedding."""
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected [n, 3, H, W]")
        n, _, H, W = x.shape
        xp = torch.zeros((n, H, W, STEM_CPAD), dtype=f16, device=self.device)
        xp[..., :3] = x.to(device=self.device, dtype=f32).permute(0, 2, 3, 1)
        return self.forward_nhwc(xp)

    __call__ = forward

    def eval(self):
        return self

    def to(self, *a, **k):
        return self


def arcface_embedding_from_images(model: IResNet, images: torch.Tensor, bbox: torch.Tensor) -> torch.Tensor:
    """images: fp32 NHWC [n, H, W, 3] in [0, 1] (the pipeline's `output_type="pt"`-style decode), bbox int32 [n, 4]
    (x0, y0, x1, y1; MTCNN's role in train_ID-Booth.py:1085-1090) -> [n, 512] ArcFace embeddings
    (crop -> bilinear 112x112 -> normalise fused in one kernel, train_ID-Booth.py:445-455)."""
    x = ops.crop_resize_norm(images.contiguous(), bbox.to(device=images.device, dtype=torch.int32).contiguous(),
                             size=112, c_pad=STEM_CPAD, half=True)
    return model.forward_nhwc(x)


def identity_loss(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """1 - cos(pred, gt) (train_ID-Booth.py:1096-1098)."""
    return 1.0 - torch.nn.functional.cosine_similarity(pred.float(), gt.float(), dim=-1)


def triplet_identity_loss(pred: torch.Tensor, positive: torch.Tensor, negative: torch.Tensor, margin: float = 1.0) -> torch.Tensor:
    """The `triplet_prior` identity loss of config 5: `TripletMarginWithDistanceLoss(distance_function=1 - cosine)` as built at
    train_ID-Booth.py:974-979 and applied at `:1133` to (predicted-x0 embedding, gt_embed[0], gt_embed[1]):
    mean(max(d(a, p) - d(a, n) + margin, 0)).  A few hundred floats: plain tensor arithmetic, no kernel."""
    cos = torch.nn.functional.cosine_similarity
    d_ap = 1.0 - cos(pred.float(), positive.float())
    d_an = 1.0 - cos(pred.float(), negative.float())
    return torch.clamp_min(margin + d_ap - d_an, 0.0).mean()


def identity_noise_level_weight(timestep, num_train_timesteps: int = 1000, timestep_loss_weighting: bool = True):
    """`(1 - t / T) ** 2` (train_ID-Booth.py:1100-1101,1129-1130): the identity term counts less at high noise levels."""
    return (1.0 - timestep / num_train_timesteps) ** 2 if timestep_loss_weighting else 1


def identity_step_tables(scheduler, timesteps, device):
    """Host part of `training_forward_identity` for per-sample timesteps: (t fp32 [B], x0 coefficient rows fp32 [B, 5]) on
    `device`.  Computing them once outside lets the rest of the chain be captured in a CUDA graph (copy new tables into
    the same tensors before each replay)."""
    ts = [int(v) for v in (timesteps.tolist() if torch.is_tensor(timesteps) else (timesteps if hasattr(timesteps, "__len__") else [timesteps]))]
    rows = []
    for tv in ts:
        c = scheduler.coef_row(None, tv, "cpu").clone()
        c[4] = 0.0   # pred_original_sample does not depend on the variance noise
        rows.append(c)
    return torch.tensor(ts, dtype=f32).to(device), torch.stack(rows).to(device)


def training_forward_identity(unet, vae, scheduler, arcface: IResNet, noisy_latents: torch.Tensor, timesteps,
                              encoder_hidden_states: torch.Tensor, bbox: torch.Tensor, context=None, tables=None):
    """The forward half of the reference's identity-loss branch (config 5; train_ID-Booth.py:1040-1046, :1081,
    :433-455, :1093): UNet(noisy, t, ctx) -> `scheduler.step(...).pred_original_sample` (the x0 estimate) -> VAE decode
    -> (x/2 + 0.5).clamp -> crop bbox -> bilinear 112x112 -> normalise -> IResNet embedding.
    Returns (model_pred [B,4,h,w], x0 latents [B,4,h,w], embeddings [B,512]).  MTCNN (third party, not in the tree)
    is replaced by the caller-supplied bbox."""
    B = noisy_latents.shape[0]
    # `tables` = identity_step_tables(...) computed by the caller: no host work inside (the chain is then graph-capturable)
    t, coefs = tables if tables is not None else identity_step_tables(scheduler, timesteps, unet.device)
    if t.numel() == 1 and B > 1:
        t, coefs = t.expand(B).contiguous(), coefs.expand(B, 5).contiguous()
    eps = unet.forward(noisy_latents, t, encoder_hidden_states=encoder_hidden_states, context=context, return_dict=False)[0]
    x0 = torch.empty_like(noisy_latents, dtype=f32)
    x_prev = torch.empty_like(x0)
    vpred = scheduler.config.prediction_type == "v_prediction"
    for i in range(B):   # per-sample timesteps (train_ID-Booth.py:1012-1018): one coefficient row each
        ops.cfg_ddpm_step(eps[i:i + 1].float().contiguous(), noisy_latents[i:i + 1].float().contiguous(), None, coefs[i],
                          guidance_scale=1.0, use_cfg=False, v_prediction=vpred, x_prev=x_prev[i:i + 1], x0_out=x0[i:i + 1])
    img = vae.decode(x0 / vae.config.scaling_factor, output_image=True)[0]          # NHWC fp32 in [0, 1]
    emb = arcface_embedding_from_images(arcface, img, bbox)
    return eps, x0, emb
