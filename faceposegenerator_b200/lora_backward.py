"""LoRA-only training backward through the frozen SD2.1 UNet.

Reference: `/root/reference/train_ID-Booth.py:1040-1046` (UNet call), `:1055-1075` (epsilon / v target, `F.mse_loss`),
`:1140-1146` (`accelerator.backward(loss)`, clip, AdamW) with the rank-4 adapters on `to_q, to_k, to_v, to_out.0` of every
attention as the only trainable tensors (`:672-678`, fp32 adapters `:779-785`).

`UNetLoRAGrad(unet)` runs the SAME forward kernels as `UNet2DConditionModel.forward` while keeping what the backward
needs (a tape of residual-stream tensors, GroupNorm sums, attention log-sum-exps), then walks the network in reverse:

* every contraction of the backward runs on the tensor-core kernels of the forward: the input gradient of a Linear /
  3x3 conv is `idb_gemm_conv` on the transposed (tap-flipped) packed weight; `dX = dY W + s (dY B) A` of an adapted
  projection is the FUSED-LoRA form of that kernel with the adapter roles swapped (down' = s B^T, up' = A^T); a stride-2
  conv's input gradient is the stride-1 conv of the zero-inserted gradient, an upsampling conv's the 2x2 sum-pool of the
  full-resolution input gradient; attention uses `idb_attention_backward` (tcgen05, recomputes P from the saved LSE);
* GroupNorm(+SiLU) / LayerNorm / GEGLU input gradients and the adapter weight gradients (`dB = dY^T (x A^T)`,
  `dA = (dY B)^T x`, tall-skinny reductions) are the bandwidth-bound kernels of `csrc/backward_kernels.cu`;
* operands of the backward GEMMs are bf16 (like the forward's), the gradient of the residual stream is fp32.

Only the adapter gradients are produced -- base weights are frozen, so no weight gradient of any base layer is formed and
the backward stops at the first adapted block (`down_blocks.0.attentions.0`).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch

from . import ops
from .packing import pack_lora

bf16, f32 = torch.bfloat16, torch.float32
HEAD_DIM = 64


# ---------------------------------------------------------------------------------------------- packed backward operands
def pack_conv_dgrad_weight(w: torch.Tensor, device=None) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> bf16 [Cin, 9 * Cout]: the input gradient of a stride-1 pad-1 conv is the conv of dY with the
    taps flipped and in / out channels swapped (tap-major, channel-minor like `pack_conv_weight`)."""
    wt = w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1)
    return wt.to(device=device, dtype=bf16).contiguous()


def pack_linear_dgrad_weight(w: torch.Tensor, device=None) -> torch.Tensor:
    """[N, K] -> bf16 [K, N] (dX = dY W)."""
    return w.t().to(device=device, dtype=bf16).contiguous()


def pack_lora_dgrad(adapters, n_in: int, seg_in: int, device=None):
    """Adapter operands of the fused-LoRA GEMM for dX = dY W + sum_seg s (dY_seg B_seg) A_seg, where segment `seg` of the
    FORWARD output (width seg_in each, dY has len(adapters) * seg_in columns) has adapter (A [r, K], B [seg_in, r], s):
    down' [16, n_seg * seg_in] holds s B_seg^T in rows seg*r .. and columns of its segment, up' [K, 64] holds A_seg^T in
    columns seg*r ..  (total rank n_seg * r <= 16)."""
    live = [a for a in adapters if a is not None]
    if not live:
        return None, None
    r = live[0][0].shape[0]
    k = live[0][0].shape[1]
    if r * len(adapters) > 16:
        raise ValueError("fused LoRA backward supports total rank <= 16")
    src = live[0][0].device
    down = torch.zeros((16, len(adapters) * seg_in), dtype=f32, device=src)
    up = torch.zeros((k, 64), dtype=f32, device=src)
    for s_, a in enumerate(adapters):
        if a is None:
            continue
        A, B, scale = a
        down[s_ * r:(s_ + 1) * r, s_ * seg_in:(s_ + 1) * seg_in] = (B.detach().float().to(src) * float(scale)).t()
        up[:, s_ * r:(s_ + 1) * r] = A.detach().float().to(src).t()
    return down.to(device=device, dtype=bf16).contiguous(), up.to(device=device, dtype=bf16).contiguous()


def pad_rows(w: torch.Tensor, rows: int) -> torch.Tensor:
    out = torch.zeros((rows, w.shape[1]), dtype=w.dtype, device=w.device)
    out[:w.shape[0]] = w
    return out


class UNetLoRAGrad:
    """Adapter gradients of a scalar loss of the UNet output, on the hand-written kernels.

        eng = UNetLoRAGrad(unet)                      # unet.set_lora(lora) already called
        eps = eng.forward(noisy, timesteps, ctx)      # same result as unet.forward; keeps the tape
        grads = eng.backward(d_loss_d_eps)            # {module_path: (dA [r, in], dB [out, r])}
    """

    def __init__(self, unet, lora: Dict[str, Tuple[torch.Tensor, torch.Tensor, float]]):
        self.u = unet
        self.dev = unet.device
        self.lora = lora
        self.rank = next(iter(lora.values()))[0].shape[0]
        self._pack_backward_weights()

    # ------------------------------------------------------------------ one-time packing of the backward operands
    def _pack_backward_weights(self):
        u, dev = self.u, self.dev
        self.res_w = {}
        self.tr_w = {}

        def unpack_conv(wp, cin, taps=9):      # forward packing [Cout, taps*Cin (+Cs)] -> ([Cout, Cin, 3, 3], shortcut [Cout, Cs] or None)
            cout = wp.shape[0]
            main = wp[:, :taps * cin].float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)
            sc = wp[:, taps * cin:].float() if wp.shape[1] > taps * cin else None
            return main, sc

        def pack_res(r):
            if id(r) in self.res_w:
                return
            w1, _ = unpack_conv(r.w1, r.cin)
            w2, sc = unpack_conv(r.w2, r.cout)
            self.res_w[id(r)] = SimpleNamespace(
                d1=pack_conv_dgrad_weight(w1, dev), d2=pack_conv_dgrad_weight(w2, dev),
                dsc=None if sc is None else sc.t().to(device=dev, dtype=bf16).contiguous())   # [Cs, Cout]: d raw = dO Wsc

        for blk in list(u.down) + list(u.up):
            for r in blk.resnets:
                pack_res(r)
        pack_res(u.mid.res0), pack_res(u.mid.res1)
        self.down_w = [None if b.down is None else pack_conv_dgrad_weight(unpack_conv(b.down[0], b.down[0].shape[1] // 9)[0], dev)
                       for b in u.down]
        self.up_w = [None if b.up is None else pack_conv_dgrad_weight(unpack_conv(b.up[0], b.up[0].shape[1] // 9)[0], dev)
                     for b in u.up]
        # conv_out (320 -> 4) input gradient = a 4 -> 320 conv of d eps: the tiny-Cin kernel with flipped, transposed fp32 weights
        wo = (u.w_conv_out[:u.out_channels].float() + u.w_conv_out[u.out_channels:2 * u.out_channels].float())      # hi + lo
        wo = wo.view(u.out_channels, 3, 3, -1).permute(0, 3, 1, 2)                                                 # [4, 320, 3, 3]
        self.w_out_dgrad = wo.flip(2, 3).permute(1, 2, 3, 0).contiguous().to(dev, f32)                              # [320, 3, 3, 4]
        for t in u.transformers:
            self.tr_w[id(t)] = SimpleNamespace(
                d_in=t.w_in.t().contiguous(), d_out=t.w_out.t().contiguous(),
                d_qkv=t.w_qkv.t().contiguous(), d_o1=t.w_o1.t().contiguous(), d_q2=t.w_q2.t().contiguous(), d_o2=t.w_o2.t().contiguous(),
                d_ff1=t.w_ff1.t().contiguous(), d_ff2=t.w_ff2.t().contiguous())
        self.update_lora(self.lora)

    def update_lora(self, lora):
        """(Re)pack only the adapter operands of the backward (a few MB), in place into persistent buffers: the transposed
        base weights are packed once.  Per adapter: s B^T and A^T slices of the fused-LoRA dgrad operands (down' [16, n_seg C],
        up' [K, 64], see `pack_lora_dgrad`) and the skinny GEMM weights of the weight gradients (A and s B^T padded to 32 rows)."""
        self.lora, dev = lora, self.dev
        r = self.rank
        for t in self.u.transformers:
            tb = t.path + ".transformer_blocks.0"
            a1, a2 = tb + ".attn1", tb + ".attn2"
            C = t.c
            w = self.tr_w[id(t)]
            if not hasattr(w, "A"):
                w.A, w.Bs = {}, {}
                for name, keys in (("l_qkv", [a1 + ".to_q", a1 + ".to_k", a1 + ".to_v"]), ("l_o1", [a1 + ".to_out.0"]),
                                   ("l_q2", [a2 + ".to_q"]), ("l_o2", [a2 + ".to_out.0"])):
                    if all(k in lora for k in keys):
                        setattr(w, name, (torch.zeros((16, len(keys) * C), dtype=bf16, device=dev), torch.zeros((C, 64), dtype=bf16, device=dev)))
                    else:    # partially adapted projection groups take the generic (allocating) packer
                        setattr(w, name, pack_lora_dgrad([lora.get(k) for k in keys], C, C, dev))
                w._groups = (("l_qkv", [a1 + ".to_q", a1 + ".to_k", a1 + ".to_v"]), ("l_o1", [a1 + ".to_out.0"]),
                             ("l_q2", [a2 + ".to_q"]), ("l_o2", [a2 + ".to_out.0"]))
            for name, keys in w._groups:
                if not all(k in lora for k in keys):
                    setattr(w, name, pack_lora_dgrad([lora.get(k) for k in keys], C, C, dev))
                    continue
                down, up = getattr(w, name)
                for s_, k in enumerate(keys):
                    A, B, scale = lora[k]
                    down[s_ * r:(s_ + 1) * r, s_ * C:(s_ + 1) * C].copy_((B.detach().float() * float(scale)).t())
                    up[:, s_ * r:(s_ + 1) * r].copy_(A.detach().t())
            for k in self._keys(t):
                if k not in lora:
                    continue
                A, B, scale = lora[k]
                if k not in w.A:
                    w.A[k] = torch.zeros((32, A.shape[1]), dtype=bf16, device=dev)
                    w.Bs[k] = torch.zeros((32, B.shape[0]), dtype=bf16, device=dev)
                w.A[k][:r].copy_(A.detach())
                w.Bs[k][:r].copy_((B.detach().float() * float(scale)).t())

    @staticmethod
    def _keys(t):
        tb = t.path + ".transformer_blocks.0"
        return [f"{tb}.{a}.{m}" for a in ("attn1", "attn2") for m in ("to_q", "to_k", "to_v", "to_out.0")]

    # ------------------------------------------------------------------ small helpers
    def _gemm(self, a0, w, **kw):
        return ops.gemm_conv(a0, w, k_splits=0, workspace=self.u._ws(), sums_pool=self.u._sums, stats_gran=self.u.stats_gran, **kw)

    def _lin_lora(self, a0, w, lo, seg_n, **kw):
        if lo[0] is None:
            return self._gemm(a0, w, **kw)
        return self._gemm(a0, w, lora_down=lo[0], lora_up=lo[1], lora_seg_n=seg_n, **kw)

    def _dgrad_lin(self, dy_b, wt, lo=None, **kw):
        """dX fp32 = dY W (+ fused adapter term); dy_b bf16 [M, N], wt bf16 [K, N]"""
        if lo is not None and lo[0] is not None:
            return self._gemm(dy_b, wt, lora_down=lo[0], lora_up=lo[1], lora_seg_n=wt.shape[0], want_f32=True, **kw)[0]
        return self._gemm(dy_b, wt, want_f32=True, **kw)[0]

    def _wgrad(self, key, x_b, dy_b, dy_col0, width, grads):
        """dB [out, r] = dY^T (x A^T), dA [r, in] = ((dY s B)^T x) for adapter `key`; x_b bf16 [M, in], dy_b bf16 [M, ...]."""
        if key not in self.lora:
            return
        t = self._cur
        r = self.rank
        scale = float(self.lora[key][2])
        # T = x A^T  [M, 32 padded] and U = dY (s B) [M, 32 padded] on the tensor cores (N = 32: one epilogue chunk)
        _, T = self._gemm(x_b, t.A[key], want_bf16=True)
        dy_seg = dy_b if (dy_col0 == 0 and dy_b.shape[1] == width) else dy_b[:, dy_col0:dy_col0 + width].contiguous()
        _, U = self._gemm(dy_seg, t.Bs[key], want_bf16=True)
        dB = ops.lora_wgrad(dy_seg, T, r, scale=scale)                            # [out, r]
        dA = ops.lora_wgrad(x_b, U, r, transpose_out=True)                        # [r, in]  (s already inside U)
        if key in grads:
            grads[key] = (grads[key][0] + dA, grads[key][1] + dB)
        else:
            grads[key] = (dA, dB)

    # ------------------------------------------------------------------ forward with tape
    def forward(self, sample, timestep, encoder_hidden_states):
        u = self.u
        x = sample.to(device=self.dev, dtype=f32).contiguous()
        B, _, H, W = x.shape
        t = timestep.to(device=self.dev, dtype=f32).reshape(-1) if torch.is_tensor(timestep) else torch.tensor([float(timestep)], device=self.dev)
        if t.numel() == 1:
            t = t.expand(B)
        t = t.contiguous()
        ctx_b = encoder_hidden_states.to(device=self.dev, dtype=bf16).reshape(-1, encoder_hidden_states.shape[-1]).contiguous()
        n_ctx = encoder_hidden_states.shape[1]
        gnws = ops.groupnorm_workspace(B, u.groups, self.dev)
        u._sums = ops.SumsPool(self.dev, capacity=B * 2 * 61440)
        temb = ops.time_embed(t, u.t_w1, u.t_b1, u.t_w2, u.t_b2, u.t_w_all, u.t_b_all)
        tape = []
        self.tape, self.ctx_b, self.n_ctx, self.B = tape, ctx_b, n_ctx, B

        h0, _, h0_st = u._gemm(ops.latent_operand(x), u.w_conv_in, mode=ops.A_3X3, bias=u.b_conv_in, want_f32=True, want_stats=True)
        h = (h0.view(B, H, W, -1), h0_st)
        skips = [h]
        for i, blk in enumerate(u.down):
            for j, r in enumerate(blk.resnets):
                h = self._resnet_fwd(r, h, None, temb, gnws)
                if blk.attns:
                    h = self._transformer_fwd(blk.attns[j], h, gnws)
                skips.append(h)
            if blk.down is not None:
                ht = h[0]
                hb = ops.cast_bf16(ht)
                o, _, o_st = u._gemm(hb, blk.down[0], mode=ops.A_3X3_S2, bias=blk.down[1], want_f32=True, want_stats=True)
                tape.append(("down", i, ht))
                h = (o.view(B, ht.shape[1] // 2, ht.shape[2] // 2, ht.shape[3]), o_st)
                skips.append(h)
        h = self._resnet_fwd(u.mid.res0, h, None, temb, gnws)
        h = self._transformer_fwd(u.mid.attn, h, gnws)
        h = self._resnet_fwd(u.mid.res1, h, None, temb, gnws)
        for i, blk in enumerate(u.up):
            for j, r in enumerate(blk.resnets):
                h = self._resnet_fwd(r, h, skips.pop(), temb, gnws)
                if blk.attns:
                    h = self._transformer_fwd(blk.attns[j], h, gnws)
            if blk.up is not None:
                ht = h[0]
                hu = ops.upsample2x(ht)
                o, _, o_st = u._gemm(hu, blk.up[0], mode=ops.A_3X3, bias=blk.up[1], want_f32=True, want_stats=True)
                tape.append(("up", i, tuple(ht.shape)))
                h = (o.view(B, hu.shape[1], hu.shape[2], ht.shape[3]), o_st)
        n, _ = ops.groupnorm(h[0], u.out_g, u.out_b, groups=u.groups, eps=u.eps, silu=True, partials=gnws, x0_stats=h[1])
        tape.append(("out", h))
        o, _ = u._gemm(n, u.w_conv_out, mode=ops.A_3X3, bias=u.b_conv_out, want_f32=True)
        o = o.view(B, H, W, 32)
        co = u.out_channels
        return (o[..., :co] + o[..., co:2 * co]).permute(0, 3, 1, 2).contiguous()

    def _resnet_fwd(self, r, hs, skip, temb, gnws):
        u = self.u
        h, h_st = hs[0], hs[1]
        sk, sk_st = skip if skip is not None else (None, None)
        B, H, W = h.shape[:3]
        n1, raw = ops.groupnorm(h, r.g1, r.b1, groups=u.groups, eps=u.eps, silu=True, x1=sk, want_raw=r.shortcut, partials=gnws,
                                x0_stats=h_st, x1_stats=sk_st)
        t1, _, t1_st = u._gemm(n1, r.w1, mode=ops.A_3X3, bias=r.bias1, rowvec=temb[:, r.temb_off:], rowvec_ld=temb.shape[1],
                               want_f32=True, want_stats=True)
        t1 = t1.view(B, H, W, r.cout)
        n2, _ = ops.groupnorm(t1, r.g2, r.b2, groups=u.groups, eps=u.eps, silu=True, partials=gnws, x0_stats=t1_st)
        if r.shortcut:
            o, _, o_st = u._gemm(n2, r.w2, mode=ops.A_3X3, a1=raw, bias=r.bias2, want_f32=True, want_stats=True)
        else:
            o, _, o_st = u._gemm(n2, r.w2, mode=ops.A_3X3, bias=r.bias2, residual=h, want_f32=True, want_stats=True)
        self.tape.append(("res", r, h, h_st, sk, sk_st, t1, t1_st))
        return o.view(B, H, W, r.cout), o_st

    def _transformer_fwd(self, t, hs, gnws):
        u = self.u
        h, h_st = hs
        B, H, W, Cc = h.shape
        M, T = B * H * W, H * W
        S = self.n_ctx
        n, _ = ops.groupnorm(h, t.gn_g, t.gn_b, groups=u.groups, eps=1e-6, silu=False, partials=gnws, x0_stats=h_st)
        x0, _ = u._gemm(n.view(M, Cc), t.w_in, bias=t.b_in, want_f32=True)
        a = ops.layernorm(x0, *t.ln[0])
        _, qkv = u._lin_lora(a, t.w_qkv, t.lora["qkv"], Cc, want_bf16=True)
        lse1 = torch.empty((B, t.heads, T), dtype=f32, device=self.dev)
        o1 = ops.attention(qkv, qkv, qkv, batch=B, heads=t.heads, t_q=T, t_kv=T, scale=HEAD_DIM ** -0.5, col0_q=0, col0_k=Cc,
                           col0_v=2 * Cc, lse=lse1)
        x1, _ = u._lin_lora(o1, t.w_o1, t.lora["o1"], Cc, bias=t.b_o1, residual=x0, want_f32=True)
        a = ops.layernorm(x1, *t.ln[1])
        _, q2 = u._lin_lora(a, t.w_q2, t.lora["q2"], Cc, want_bf16=True)
        kv = u._context_kv(t, self.ctx_b)
        lse2 = torch.empty((B, t.heads, T), dtype=f32, device=self.dev)
        o2 = ops.attention(q2, kv, kv, batch=B, heads=t.heads, t_q=T, t_kv=S, scale=HEAD_DIM ** -0.5, col0_q=0, col0_k=0, col0_v=Cc,
                           lse=lse2)
        x2, _ = u._lin_lora(o2, t.w_o2, t.lora["o2"], Cc, bias=t.b_o2, residual=x1, want_f32=True)
        a = ops.layernorm(x2, *t.ln[2])
        _, g = u._gemm(a, t.w_ff1, bias=t.b_ff1, geglu=True, want_bf16=True)
        _, x3 = u._gemm(g, t.w_ff2, bias=t.b_ff2, residual=x2, want_bf16=True)
        out, _, out_st = u._gemm(x3, t.w_out, bias=t.b_out, residual=h.view(M, Cc), want_f32=True, want_stats=True, stats_hw=T)
        self.tape.append(("tr", t, h, h_st, x0, x1, x2, qkv, o1, lse1, q2, kv, o2, lse2))
        return out.view(B, H, W, Cc), out_st

    # ------------------------------------------------------------------ backward
    def backward(self, d_eps: torch.Tensor) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
        """d_eps: gradient of the loss with respect to the UNet output, [B, 4, H, W].  Returns {path: (dA, dB)} (fp32)."""
        u = self.u
        grads: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        tape = list(self.tape)
        d_eps = d_eps.to(device=self.dev, dtype=f32).contiguous()
        B, _, H, W = d_eps.shape
        kind, h_out = tape.pop()
        assert kind == "out"
        # conv_out input gradient (4 -> 320 tiny-Cin conv with flipped / transposed weights), then conv_norm_out + SiLU
        dn, _ = ops.conv3x3_small_cin(d_eps, self.w_out_dgrad, None, nchw=True)
        hx, hst = h_out
        Cc = hx.shape[-1]
        if hst is None or hst.dtype != torch.int64:
            raise NotImplementedError("the LoRA backward needs the forward's per-image GroupNorm sums: latent height / width must be "
                                      "powers of two (64 x 64 latents = 512 x 512 training images, train_ID-Booth.py resolution)")
        stats = ops.group_stats_from_sums(hst, Cc, hx.shape[1] * hx.shape[2], u.groups, u.eps)
        dh, _ = ops.groupnorm_backward(dn, hx, u.out_g, u.out_b, stats, groups=u.groups, silu=True)
        # gradients of skip tensors, keyed by the tensor object they belong to (summed into the producer's gradient)
        dskip: Dict[int, torch.Tensor] = {}
        first_adapted = u.transformers[0]
        while tape:
            rec = tape.pop()
            if rec[0] == "up":
                _, i, shp = rec
                # Upsample2D: o = conv3x3(nearest2x(h)) -> d up = dgrad conv at full resolution, d h = 2x2 sum-pool
                d_up = self._gemm(ops.cast_bf16(dh), self.up_w[i], mode=ops.A_3X3, want_f32=True)[0].view(dh.shape[0], dh.shape[1], dh.shape[2], -1)
                dh = ops.sumpool2x(d_up)
            elif rec[0] == "down":
                _, i, ht = rec
                # Downsample2D (stride-2 conv): input gradient = stride-1 conv of the zero-inserted gradient with flipped weights
                z = ops.zero_insert2x(ops.cast_bf16(dh))
                dh = self._with_skip(ht, self._gemm(z, self.down_w[i], mode=ops.A_3X3, want_f32=True)[0].view(ht.shape), dskip)
            elif rec[0] == "res":
                dh = self._resnet_bwd(rec, dh, dskip)
            elif rec[0] == "tr":
                dh = self._with_skip(rec[2], self._transformer_bwd(rec, dh, grads), dskip)
                if rec[1] is first_adapted:
                    break          # nothing trainable lies further upstream
        return grads

    @staticmethod
    def _with_skip(x, dx, dskip):
        """`dx` = gradient of tensor `x` from the record just processed; if `x` was also stored as a skip connection, add the
        gradient its up-block consumer produced (every residual-stream tensor has at most these two consumers)."""
        extra = dskip.pop(id(x), None)
        return dx if extra is None else dx + extra

    def _resnet_bwd(self, rec, do, dskip):
        u = self.u
        _, r, h, h_st, sk, sk_st, t1, t1_st = rec
        w = self.res_w[id(r)]
        B, H, W = h.shape[:3]
        cin0, cin1 = h.shape[-1], (0 if sk is None else sk.shape[-1])
        do_b = ops.cast_bf16(do)
        dn2 = self._gemm(do_b, w.d2, mode=ops.A_3X3, want_f32=True)[0].view(B, H, W, r.cout)
        st2 = ops.group_stats_from_sums(t1_st, r.cout, H * W, u.groups, u.eps)
        dt1, _ = ops.groupnorm_backward(dn2, t1, r.g2, r.b2, st2, groups=u.groups, silu=True)
        dn1 = self._gemm(ops.cast_bf16(dt1), w.d1, mode=ops.A_3X3, want_f32=True)[0].view(B, H, W, cin0 + cin1)
        st1 = ops.group_stats_from_sums(h_st, cin0, H * W, u.groups, u.eps, sk_st, cin1)
        # the shortcut path: identity (dh += do) or the 1x1 conv_shortcut over [h | skip] (d raw = dO Wsc)
        if r.shortcut:
            draw = self._gemm(do_b.view(B * H * W, r.cout), w.dsc, want_f32=True)[0].view(B, H, W, cin0 + cin1)
            dh0 = draw[..., :cin0].contiguous()
            dsk = draw[..., cin0:].contiguous() if sk is not None else None
        else:
            dh0, dsk = do.clone(), None
        dh0, dsk = ops.groupnorm_backward(dn1, h, r.g1, r.b1, st1, groups=u.groups, silu=True, x1=sk, dx0=dh0, dx1=dsk,
                                          add0=True, add1=dsk is not None)
        if sk is not None:
            dskip[id(sk)] = dsk if id(sk) not in dskip else dskip[id(sk)] + dsk
        return self._with_skip(h, dh0, dskip)

    def _transformer_bwd(self, rec, dout, grads):
        u = self.u
        _, t, h, h_st, x0, x1, x2, qkv, o1, lse1, q2, kv, o2, lse2 = rec
        w = self.tr_w[id(t)]
        self._cur = w
        tb = t.path + ".transformer_blocks.0"
        a1k, a2k = tb + ".attn1", tb + ".attn2"
        B, H, W, Cc = h.shape
        M, T, S = B * H * W, H * W, self.n_ctx
        scale = HEAD_DIM ** -0.5
        dout2 = dout.reshape(M, Cc)
        # proj_out (+ h residual)
        dx = self._dgrad_lin(ops.cast_bf16(dout2), w.d_out)                    # d x3 = d x2 (residual) so far
        # feed-forward: x3 = g W2^T + b + x2, g = GEGLU(LN3(x2) W1^T + b)
        dxb = ops.cast_bf16(dx)
        _, dg = self._gemm(dxb, w.d_ff2, want_bf16=True)                       # [M, 4C] bf16
        a3 = ops.layernorm(x2, *t.ln[2])
        _, pre = self._gemm(a3, t.w_ff1, bias=t.b_ff1, want_bf16=True)         # recomputed pre-activation, interleaved [M, 8C]
        du = ops.geglu_backward(dg, pre)
        da3 = self._dgrad_lin(du, w.d_ff1)
        ops.layernorm_backward(da3, x2, t.ln[2][0], dx=dx, add=True)           # dx = d x2
        # cross-attention: x2 = o2 Wo2^T + b (+LoRA) + x1
        dxb = ops.cast_bf16(dx)
        do2 = self._dgrad_lin(dxb, w.d_o2, w.l_o2)
        self._wgrad(a2k + ".to_out.0", o2, dxb, 0, Cc, grads)
        do2b = ops.cast_bf16(do2)
        dq2, dk2, dv2 = ops.attention_backward(q2, kv, kv, o2, do2b, lse2, batch=B, heads=t.heads, t_q=T, t_kv=S, scale=scale,
                                               col0_q=0, col0_k=0, col0_v=Cc)
        dq2b = ops.cast_bf16(dq2)
        a2 = ops.layernorm(x1, *t.ln[1])
        da2 = self._dgrad_lin(dq2b, w.d_q2, w.l_q2)
        self._wgrad(a2k + ".to_q", a2, dq2b, 0, Cc, grads)
        self._wgrad(a2k + ".to_k", self.ctx_b, dk2, 0, Cc, grads)
        self._wgrad(a2k + ".to_v", self.ctx_b, dv2, 0, Cc, grads)
        ops.layernorm_backward(da2, x1, t.ln[1][0], dx=dx, add=True)           # dx = d x1
        # self-attention: x1 = o1 Wo1^T + b (+LoRA) + x0
        dxb = ops.cast_bf16(dx)
        do1 = self._dgrad_lin(dxb, w.d_o1, w.l_o1)
        self._wgrad(a1k + ".to_out.0", o1, dxb, 0, Cc, grads)
        do1b = ops.cast_bf16(do1)
        dqkv_f = torch.zeros((M, Cc), dtype=f32, device=self.dev)
        dqkv_b = torch.empty((M, 3 * Cc), dtype=bf16, device=self.dev)
        ops.attention_backward(qkv, qkv, qkv, o1, do1b, lse1, batch=B, heads=t.heads, t_q=T, t_kv=T, scale=scale, col0_q=0, col0_k=Cc,
                               col0_v=2 * Cc, dq=dqkv_f, dk=dqkv_b, dv=dqkv_b, col0_dk=Cc, col0_dv=2 * Cc)
        dqkv_b[:, :Cc] = dqkv_f.to(bf16)
        a1 = ops.layernorm(x0, *t.ln[0])
        da1 = self._dgrad_lin(dqkv_b, w.d_qkv, w.l_qkv)
        self._wgrad(a1k + ".to_q", a1, dqkv_b, 0, Cc, grads)
        self._wgrad(a1k + ".to_k", a1, dqkv_b, Cc, Cc, grads)
        self._wgrad(a1k + ".to_v", a1, dqkv_b, 2 * Cc, Cc, grads)
        ops.layernorm_backward(da1, x0, t.ln[0][0], dx=dx, add=True)           # dx = d x0
        # proj_in and the block's GroupNorm; + the residual h
        dn = self._dgrad_lin(ops.cast_bf16(dx), w.d_in)
        stats = ops.group_stats_from_sums(h_st, Cc, T, u.groups, 1e-6)
        dh = dout.clone()
        ops.groupnorm_backward(dn.view(B, H, W, Cc), h, t.gn_g, t.gn_b, stats, groups=u.groups, silu=False, dx0=dh, add0=True)
        return dh


class LoRATrainer:
    """The optimisation step of `/root/reference/train_ID-Booth.py:1012-1146` for the denoising loss (`which_loss == ""`:
    `instance_loss = F.mse_loss(model_pred, target)`), LoRA-only: `add_noise` (`:1018`) -> UNet (`:1040-1046`) -> epsilon /
    v target (`:1055-1058`) -> MSE (`:1137-1138`) -> backward (`:1140`) -> `clip_grad_norm_(max_grad_norm)` (`:1143`) ->
    AdamW (`:1144`; lr 1e-4, betas (0.9, 0.999), weight decay 1e-2, eps 1e-8: `configs/config_train_SD21.py:58-66`).

    Forward and backward run on the hand-written kernels (`UNetLoRAGrad`); the 128 fp32 adapter factors (829,952 values)
    and their AdamW state are ordinary torch tensors updated by `torch.optim.AdamW`, then re-installed in place into the
    UNet's packed adapter buffers (`set_lora`).  The identity / triplet losses of the reference additionally differentiate
    through the VAE decoder and the ArcFace backbone (`:1081-1133`); that backward is not built (forward only:
    `iresnet.training_forward_identity`)."""

    def __init__(self, unet, lora, scheduler, lr: float = 1e-4, betas=(0.9, 0.999), weight_decay: float = 1e-2, eps: float = 1e-8,
                 max_grad_norm: float = 1.0, use_cuda_graph: Optional[bool] = None):
        self.unet, self.scheduler, self.max_grad_norm = unet, scheduler, max_grad_norm
        dev = unet.device
        # add_noise -> forward with tape -> loss -> backward as ONE CUDA-graph launch per batch geometry (about 1,900 kernels:
        # eager, the host launch rate adds ~20 ms to a 47 ms step); the adapters are re-installed in place, so the graph
        # survives the optimiser steps.  IDB_CUDA_GRAPH=0 / use_cuda_graph=False: eager.
        import os
        self.use_cuda_graph = (os.environ.get("IDB_CUDA_GRAPH", "1") != "0") if use_cuda_graph is None else bool(use_cuda_graph)
        self._graphs = {}
        self._acp = scheduler.alphas_cumprod.to(device=dev, dtype=f32)
        self.params = {k: (torch.nn.Parameter(d.detach().to(dev, f32).clone()), torch.nn.Parameter(u.detach().to(dev, f32).clone()), float(s))
                       for k, (d, u, s) in lora.items()}
        self.opt = torch.optim.AdamW([p for d, u, _ in self.params.values() for p in (d, u)], lr=lr, betas=betas,
                                     weight_decay=weight_decay, eps=eps)
        self._install()

    def lora(self):
        return {k: (d.detach(), u.detach(), s) for k, (d, u, s) in self.params.items()}

    def _install(self):
        """Re-install the updated adapters: in place into the UNet's persistent packed buffers (packed on the GPU) and into
        the backward's adapter operands; base weights (forward and transposed backward packing) are untouched."""
        lora = self.lora()
        self.unet.set_lora(lora)
        if getattr(self, "engine", None) is None:
            self.engine = UNetLoRAGrad(self.unet, lora)
        else:
            self.engine.update_lora(lora)

    def _loss_and_grads(self, latents, noise, timesteps, ctx):
        """Device-only (capturable): `scheduler.add_noise` / `get_velocity` arithmetic on the device copy of alphas_cumprod,
        forward with tape, MSE, backward.  Returns (loss scalar tensor, {adapter: (dA, dB)})."""
        acp = self._acp[timesteps]
        sa, sb = acp.sqrt().view(-1, 1, 1, 1), (1.0 - acp).sqrt().view(-1, 1, 1, 1)
        noisy = sa * latents + sb * noise
        pred = self.engine.forward(noisy, timesteps.to(f32), ctx)
        target = (sa * noise - sb * latents) if self.scheduler.config.prediction_type == "v_prediction" else noise
        diff = pred - target
        loss = (diff * diff).mean()
        return loss, self.engine.backward(diff * (2.0 / diff.numel()))

    @torch.no_grad()
    def step(self, latents, noise, timesteps, encoder_hidden_states):
        """One training step on clean latents [B, 4, h, w] (already scaled by the VAE factor), noise, long timesteps [B],
        text states [B, 77, 1024].  Returns (loss, gradient norm before clipping)."""
        dev = self.unet.device
        latents, noise = latents.to(dev, f32), noise.to(dev, f32)
        timesteps = timesteps.to(dev).long()
        ctx = encoder_hidden_states.to(dev)
        if not self.use_cuda_graph:
            loss, grads = self._loss_and_grads(latents, noise, timesteps, ctx)
        else:
            key = (tuple(latents.shape), tuple(ctx.shape), ctx.dtype, self.scheduler.config.prediction_type)
            st = self._graphs.get(key)
            if st is None:
                with torch.cuda.device(dev):
                    bufs = (latents.clone(), noise.clone(), timesteps.clone(), ctx.clone())
                    self._loss_and_grads(*bufs)                     # warm-up outside capture (lazy workspaces / kernel attributes)
                    torch.cuda.synchronize(dev)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        out = self._loss_and_grads(*bufs)
                if len(self._graphs) >= 2:
                    self._graphs.pop(next(iter(self._graphs)))
                st = self._graphs[key] = (g, bufs, out)
            g, bufs, (loss, grads) = st
            for b_, v in zip(bufs, (latents, noise, timesteps, ctx)):
                b_.copy_(v)
            g.replay()
        for k, (d, u, _) in self.params.items():
            d.grad, u.grad = grads[k][0], grads[k][1]
        norm = torch.nn.utils.clip_grad_norm_([p for d, u, _ in self.params.values() for p in (d, u)], self.max_grad_norm)   # (foreach)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        self._install()
        return float(loss), float(norm)
