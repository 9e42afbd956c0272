"""`AutoencoderKL` (diffusers 0.32.2, SD2.1-base VAE) on the same sm_100a kernels as the UNet:
`vae.decode(z).sample` / `.config.scaling_factor` (`/root/reference/train_ID-Booth.py:410-412,435-437`; pipeline tail
behind `inference_ID-Booth.py:138`) and `vae.encode(x).latent_dist.sample()` (`train_ID-Booth.py:1001-1002`).
NHWC, fp32 residual stream, bf16 GEMM/conv operands.

Mid-block attention (1 head, d = 512, 4096 tokens) is expressed with the tcgen05 GEMM:
S = Q K^T (fp32), row softmax, O = P V with V^T produced directly by a swapped-operand GEMM
(V^T = W_v X^T); the to_v bias is folded into to_out's bias (softmax rows sum to 1).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import ops
from .packing import pack_conv_weight, pack_edge_conv_weight, pack_upsample_phase_weights
from .weights import VAE_CONFIG, random_state_dict, vae_decoder_manifest, vae_encoder_manifest

bf16, f32 = torch.bfloat16, torch.float32


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


class DiagonalGaussianDistribution:
    """`latent_dist` of `AutoencoderKL.encode` (diffusers): moments [n, 2*lc, h, w] = (mean | logvar)."""

    def __init__(self, moments: torch.Tensor):
        self.mean, logvar = moments.chunk(2, dim=1)
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator=None) -> torch.Tensor:
        from .scheduler import randn_tensor
        noise = randn_tensor(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


class EncoderOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist

    def __getitem__(self, i):
        return (self.latent_dist,)[i]


class AutoencoderKL:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config: dict = VAE_CONFIG, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("AutoencoderKL runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.cfg = dict(config)
        self.config = SimpleNamespace(**self.cfg)
        self.dtype = bf16
        self.groups = self.cfg["norm_num_groups"]
        self.eps = 1e-6
        self._workspace = None
        self._sums = None
        import math
        self.stats_gran = max(1, math.gcd(*config["block_out_channels"]) // config["norm_num_groups"])   # see unet.py (SD2.1 VAE: 4)
        self._pack(state_dict)

    @classmethod
    def from_random(cls, seed: int = 0, config: dict = VAE_CONFIG, device="cuda:0", with_encoder: bool = False):
        m = vae_decoder_manifest(config) + (vae_encoder_manifest(config) if with_encoder else [])
        return cls(random_state_dict(m, seed), config, device)

    def _dev(self, t, dtype=f32):
        return t.to(device=self.device, dtype=dtype).contiguous()

    def _pack_resnet(self, sd, p):
        r = SimpleNamespace()
        w1 = sd[p + ".conv1.weight"]
        r.cout, r.cin = w1.shape[0], w1.shape[1]
        r.g1, r.b1 = self._dev(sd[p + ".norm1.weight"]), self._dev(sd[p + ".norm1.bias"])
        r.w1, r.bias1 = pack_conv_weight(w1, device=self.device), self._dev(sd[p + ".conv1.bias"])
        r.g2, r.b2 = self._dev(sd[p + ".norm2.weight"]), self._dev(sd[p + ".norm2.bias"])
        sc = sd.get(p + ".conv_shortcut.weight")
        r.shortcut = sc is not None
        r.w2 = pack_conv_weight(sd[p + ".conv2.weight"], shortcut=sc, device=self.device)
        b2 = sd[p + ".conv2.bias"].float()
        if r.shortcut:
            b2 = b2 + sd[p + ".conv_shortcut.bias"].float()
        r.bias2 = self._dev(b2)
        return r

    def _pack_attn(self, sd, a):
        at = SimpleNamespace()
        at.g, at.b = self._dev(sd[a + ".group_norm.weight"]), self._dev(sd[a + ".group_norm.bias"])
        at.wq, at.bq = self._dev(sd[a + ".to_q.weight"], bf16), self._dev(sd[a + ".to_q.bias"])
        at.wk, at.bk = self._dev(sd[a + ".to_k.weight"], bf16), self._dev(sd[a + ".to_k.bias"])
        at.wv = self._dev(sd[a + ".to_v.weight"], bf16)
        wo = sd[a + ".to_out.0.weight"].float()
        at.wo = self._dev(wo, bf16)
        # out = W_o (P V0 + b_v) + b_o  (softmax rows sum to one)
        at.bo = self._dev(sd[a + ".to_out.0.bias"].float() + wo @ sd[a + ".to_v.bias"].float())
        at.c = wo.shape[0]
        return at

    def _pack_encoder(self, sd):
        """Encoder + quant_conv (optional: only checkpoints / manifests that carry them)."""
        cfg = self.cfg
        ch = tuple(cfg["block_out_channels"])
        wi = sd["encoder.conv_in.weight"].float()                         # [128, 3, 3, 3] -> Cin padded to 4
        wp = torch.zeros(wi.shape[0], 4, 3, 3)
        wp[:, :3] = wi
        self.e_w_in = pack_edge_conv_weight(wp, self.device)
        self.e_b_in = self._dev(sd["encoder.conv_in.bias"])
        self.e_down = []
        for i in range(len(ch)):
            blk = SimpleNamespace(resnets=[self._pack_resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}")
                                           for j in range(cfg["layers_per_block"])], down=None)
            if i < len(ch) - 1:
                q = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                blk.down = (pack_conv_weight(sd[q + ".weight"], device=self.device), self._dev(sd[q + ".bias"]))
            self.e_down.append(blk)
        self.e_mid0 = self._pack_resnet(sd, "encoder.mid_block.resnets.0")
        self.e_attn = self._pack_attn(sd, "encoder.mid_block.attentions.0")
        self.e_mid1 = self._pack_resnet(sd, "encoder.mid_block.resnets.1")
        self.e_out_g, self.e_out_b = self._dev(sd["encoder.conv_norm_out.weight"]), self._dev(sd["encoder.conv_norm_out.bias"])
        # conv_out (512 -> 8) with the 1x1 quant_conv folded in (both linear), Cout padded to one 32-column chunk;
        # rows 8-15 carry the bf16 rounding residual of the folded weights (w = hi + lo)
        q = sd["quant_conv.weight"].double().reshape(sd["quant_conv.weight"].shape[0], -1)          # [8, 8]
        wo = sd["encoder.conv_out.weight"].double().permute(0, 2, 3, 1).reshape(q.shape[1], -1)      # [8, 9*512]
        wf = (q @ wo).float()
        bfold = (q @ sd["encoder.conv_out.bias"].double() + sd["quant_conv.bias"].double()).float()
        hi = wf.to(bf16)
        co = wf.shape[0]
        w32 = torch.zeros((32, wf.shape[1]), dtype=bf16)
        w32[:co], w32[co:2 * co] = hi, (wf - hi.float()).to(bf16)
        self.e_w_out = w32.to(self.device).contiguous()
        self.e_b_out = torch.zeros(32, dtype=f32, device=self.device)
        self.e_b_out[:co] = bfold.to(self.device)
        self.e_moments = co

    def _pack(self, sd):
        sd = dict(sd)
        for a in ("decoder.mid_block.attentions.0", "encoder.mid_block.attentions.0"):
            for old, new in (("query", "to_q"), ("key", "to_k"), ("value", "to_v"), ("proj_attn", "to_out.0")):
                for suf in (".weight", ".bias"):  # legacy checkpoint spelling (App. A.4)
                    key = f"{a}.{old}{suf}"
                    if key in sd:
                        t = sd.pop(key)
                        sd[f"{a}.{new}{suf}"] = t.reshape(t.shape[0], t.shape[1]) if suf == ".weight" else t
        self.has_encoder = "encoder.conv_in.weight" in sd
        if self.has_encoder:
            self._pack_encoder(sd)
        a = "decoder.mid_block.attentions.0"
        self.pq_w = self._dev(sd["post_quant_conv.weight"].reshape(4, 4))
        self.pq_b = self._dev(sd["post_quant_conv.bias"])
        self.w_in = pack_edge_conv_weight(sd["decoder.conv_in.weight"], self.device)
        self.b_in = self._dev(sd["decoder.conv_in.bias"])
        self.mid0 = self._pack_resnet(sd, "decoder.mid_block.resnets.0")
        self.mid1 = self._pack_resnet(sd, "decoder.mid_block.resnets.1")
        self.attn = self._pack_attn(sd, a)
        n = len(self.cfg["block_out_channels"])
        self.up = []
        for i in range(n):
            blk = SimpleNamespace(resnets=[self._pack_resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}")
                                           for j in range(self.cfg["layers_per_block"] + 1)], up=None)
            if i < n - 1:
                q = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                blk.up = (pack_conv_weight(sd[q + ".weight"], device=self.device), self._dev(sd[q + ".bias"]),
                          *reversed(pack_upsample_phase_weights(sd[q + ".weight"], device=self.device, stacked=True)))   # [2]: per-phase views, [3]: stacked
            self.up.append(blk)
        self.out_g, self.out_b = self._dev(sd["decoder.conv_norm_out.weight"]), self._dev(sd["decoder.conv_norm_out.bias"])
        self.w_out = pack_edge_conv_weight(sd["decoder.conv_out.weight"], self.device)
        self.b_out = self._dev(sd["decoder.conv_out.bias"])

    def _ws(self):
        if self._workspace is None:
            self._workspace = torch.empty((96 << 20) // 4, dtype=f32, device=self.device)
        return self._workspace

    def _gemm(self, a0, w, **kw):
        return ops.gemm_conv(a0, w, k_splits=0, workspace=self._ws(), sums_pool=self._sums, stats_gran=self.stats_gran, **kw)

    # stream value = (fp32 NHWC tensor, row-block channel statistics or None), see unet.py
    def _resnet(self, r, hs, gnws):
        h, h_st = hs[0], hs[1]
        h_ph = hs[2] if len(hs) > 2 else 0          # 4: statistics written by the four phased upsample GEMMs
        B, H, W = h.shape[:3]
        n1, raw = ops.groupnorm(h, r.g1, r.b1, groups=self.groups, eps=self.eps, silu=True, want_raw=r.shortcut,
                                partials=gnws, x0_stats=h_st, x0_stats_phases=h_ph)
        t1, _, t1_st = self._gemm(n1, r.w1, mode=ops.A_3X3, bias=r.bias1, want_f32=True, want_stats=True)
        n2, _ = ops.groupnorm(t1.view(B, H, W, r.cout), r.g2, r.b2, groups=self.groups, eps=self.eps, silu=True,
                              partials=gnws, x0_stats=t1_st)
        if r.shortcut:
            o, _, o_st = self._gemm(n2, r.w2, mode=ops.A_3X3, a1=raw, bias=r.bias2, want_f32=True, want_stats=True)
        else:
            o, _, o_st = self._gemm(n2, r.w2, mode=ops.A_3X3, bias=r.bias2, residual=h, want_f32=True, want_stats=True)
        return o.view(B, H, W, r.cout), o_st

    def _attention(self, hs, gnws, at=None):
        h, h_st = hs
        at = at or self.attn
        B, H, W, Cc = h.shape
        T = H * W
        n, _ = ops.groupnorm(h, at.g, at.b, groups=self.groups, eps=self.eps, silu=False, partials=gnws, x0_stats=h_st)
        n2 = n.view(B * T, Cc)
        _, q = self._gemm(n2, at.wq, bias=at.bq, want_bf16=True)
        _, k = self._gemm(n2, at.wk, bias=at.bk, want_bf16=True)
        o = torch.empty((B * T, Cc), dtype=bf16, device=h.device)
        s = torch.empty((T, T), dtype=f32, device=h.device)
        p = torch.empty((T, T), dtype=bf16, device=h.device)
        vt = torch.empty((Cc, T), dtype=bf16, device=h.device)
        for b in range(B):
            sl = slice(b * T, (b + 1) * T)
            self._gemm(at.wv, n2[sl], out_bf16=vt)             # V^T = W_v X^T  [C, T]
            self._gemm(q[sl], k[sl], out_f32=s)                # S = Q K^T      [T, T]
            ops.softmax_rows(s, Cc ** -0.5, out=p)
            self._gemm(p, vt, out_bf16=o[sl])                  # O = P V        [T, C]
        out, _, out_st = self._gemm(o, at.wo, bias=at.bo, residual=h.view(B * T, Cc), want_f32=True, want_stats=True,
                                    stats_hw=T)
        return out.view(B, H, W, Cc), out_st

    def decode(self, z, return_dict: bool = True, generator=None, output_image: bool = False,
               taps: Optional[dict] = None):
        """z: [n, 4, h, w] latents ALREADY divided by scaling_factor -> [n, 3, 8h, 8w] in ~[-1, 1].
        `output_image=True` instead returns the post-processed NHWC fp32 image in [0, 1]
        (VaeImageProcessor.postprocess "np" fused into the conv_out kernel)."""
        in_dtype = z.dtype
        z = z.to(device=self.device, dtype=f32).contiguous()
        B = z.shape[0]
        gnws = ops.groupnorm_workspace(B, self.groups, self.device)
        self._sums = ops.SumsPool(self.device, capacity=B * 2 * 16384)
        x = ops.vae_latent_prep(z, self.pq_w, self.pq_b, 1.0)
        h0, _ = ops.conv3x3_small_cin(x, self.w_in, self.b_in, nchw=False)
        h = (h0, None)
        h = self._resnet(self.mid0, h, gnws)
        h = self._attention(h, gnws)
        h = self._resnet(self.mid1, h, gnws)
        if taps is not None:
            taps["mid"] = h[0].permute(0, 3, 1, 2).clone()
        for i, blk in enumerate(self.up):
            for r in blk.resnets:
                h = self._resnet(r, h, gnws)
            if blk.up is not None:
                ht = h[0]
                Hl, Wl, Cu = ht.shape[1], ht.shape[2], ht.shape[3]
                if ops.epilogue_stats_supported(B, Hl, Wl):   # Upsample2D as four 2x2 convs on the low-resolution tensor (see unet.py)
                    xb = ops.cast_bf16(ht)
                    o = torch.empty((B, 2 * Hl, 2 * Wl, Cu), dtype=f32, device=self.device)
                    o_st = torch.empty((4, B * Hl * Wl // 32, Cu, 2), dtype=f32, device=self.device)
                    o_sums = self._sums.take(B, Cu // self.stats_gran) if ops.image_sums_supported(B, Hl * Wl, Cu, phased=True) else None
                    if ops.PHASES4_ON and (Cu % 160 == 0 or Cu % 128 == 0):
                        # all four parity classes in ONE launch (phase weights stacked on N: 4x the tiles, one fixed cost)
                        self._gemm(xb, blk.up[3], mode=ops.A_2X2, bias=blk.up[1], out_f32=o, stats=o_st, sums=o_sums, phases4=True)
                    else:
                        for a in range(2):
                            for c in range(2):
                                self._gemm(xb, blk.up[2][a][c], mode=ops.A_2X2, bias=blk.up[1], out_f32=o, stats=o_st,
                                           sums=o_sums, tap_off=(a - 1, c - 1), out_phase=(a, c))
                    h = (o, o_sums, 0) if o_sums is not None else (o, o_st, 4)
                else:
                    hu = ops.upsample2x(ht)
                    o, _, o_st = self._gemm(hu, blk.up[0], mode=ops.A_3X3, bias=blk.up[1], want_f32=True, want_stats=True)
                    h = (o.view(B, hu.shape[1], hu.shape[2], Cu), o_st)
            if taps is not None:
                taps[f"up{i}"] = h[0].permute(0, 3, 1, 2).clone()
        n, _ = ops.groupnorm(h[0], self.out_g, self.out_b, groups=self.groups, eps=self.eps, silu=True, partials=gnws,
                             x0_stats=h[1])
        img = ops.conv3x3_small_cout(n, self.w_out, self.b_out, postprocess=output_image)
        if not output_image and in_dtype != f32:
            img = img.to(in_dtype)
        return DecoderOutput(img) if return_dict else (img,)

    # ------------------------------------------------------------------ encode (train_ID-Booth.py:1001-1002)
    def encode(self, x, return_dict: bool = True):
        """x: [n, 3, H, W] in [-1, 1] -> `.latent_dist` (DiagonalGaussianDistribution over [n, 4, H/8, W/8])."""
        if not self.has_encoder:
            raise RuntimeError("this AutoencoderKL was built without encoder weights")
        x = x.to(device=self.device, dtype=f32)
        B, cin, H, W = x.shape
        if cin != 3 or H % 8 or W % 8:
            raise ValueError("expected [n, 3, H, W] with H, W multiples of 8")
        gnws = ops.groupnorm_workspace(B, self.groups, self.device)
        self._sums = ops.SumsPool(self.device, capacity=B * 2 * 16384)
        x4 = torch.zeros((B, H, W, 4), dtype=f32, device=self.device)
        x4[..., :3] = x.permute(0, 2, 3, 1)
        h0, _ = ops.conv3x3_small_cin(x4, self.e_w_in, self.e_b_in, nchw=False)
        h = (h0, None)
        for blk in self.e_down:
            for r in blk.resnets:
                h = self._resnet(r, h, gnws)
            if blk.down is not None:   # Downsample2D: F.pad(0,1,0,1) + conv3x3 stride 2 -> the kernel's asymmetric stride-2 view
                ht = h[0]
                o, _, o_st = self._gemm(ops.cast_bf16(ht), blk.down[0], mode=ops.A_3X3_S2_ASYM, bias=blk.down[1],
                                        want_f32=True, want_stats=True)
                h = (o.view(B, ht.shape[1] // 2, ht.shape[2] // 2, ht.shape[3]), o_st)
        h = self._resnet(self.e_mid0, h, gnws)
        h = self._attention(h, gnws, self.e_attn)
        h = self._resnet(self.e_mid1, h, gnws)
        n, _ = ops.groupnorm(h[0], self.e_out_g, self.e_out_b, groups=self.groups, eps=self.eps, silu=True, partials=gnws,
                             x0_stats=h[1])
        o, _ = self._gemm(n, self.e_w_out, mode=ops.A_3X3, bias=self.e_b_out, want_f32=True)
        co = self.e_moments
        o = o.view(B, H // 8, W // 8, 32)
        moments = (o[..., :co] + o[..., co:2 * co]).permute(0, 3, 1, 2).contiguous()     # hi + lo weight halves; NHWC -> NCHW
        dist = DiagonalGaussianDistribution(moments)
        return EncoderOutput(dist) if return_dict else (dist,)

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def requires_grad_(self, flag: bool = False):
        return self
