"""Host side of `UNet2DConditionModel` (diffusers 0.32.2 semantics, SD2.1-base config) on the
hand-written sm_100a kernels.  Same call surface as the reference uses
(`/root/reference/train_ID-Booth.py:1040-1046`): `unet(sample, timestep,
encoder_hidden_states, class_labels=None, return_dict=False)[0]`, `.config.in_channels`.

Data layout in HBM: activations NHWC; the residual stream is fp32, every GEMM / conv operand
is a bf16 tensor produced by the preceding norm (GroupNorm+SiLU / LayerNorm) or epilogue.
`torch.cat([h, skip], 1)` is never materialised in fp32: GroupNorm reads both sources and
writes the concatenated bf16 operand(s).  ResnetBlock2D.conv_shortcut is folded into conv2's
GEMM as extra K columns; the 22 `time_emb_proj` layers run as one batched kernel; LoRA
adapters are fused (unmerged) into the q/k/v/out projection GEMMs.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from . import ops
from .packing import interleave_geglu, pack_conv_weight, pack_lora, pack_upsample_phase_weights
from .weights import UNET_CONFIG, random_state_dict, unet_manifest

bf16, f32 = torch.bfloat16, torch.float32
HEAD_DIM = 64
WORKSPACE_BYTES = 96 << 20


class UNetOutput:
    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


class _Resnet:
    __slots__ = ("cin", "cout", "g1", "b1", "w1", "bias1", "temb_off", "g2", "b2", "w2", "bias2", "shortcut")


class _Transformer:
    pass


class UNet2DConditionModel:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config: dict = UNET_CONFIG, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("UNet2DConditionModel runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.cfg = dict(config)
        self.config = SimpleNamespace(**self.cfg)
        self.dtype = bf16
        self.groups = self.cfg["norm_num_groups"]
        self.eps = self.cfg["norm_eps"]
        self._lora_version = 0
        self._sums = None
        # channel granule of the per-image GroupNorm sums the GEMM epilogues accumulate: it must divide the group size of
        # every GroupNorm (plain and skip-concatenated) and every concat offset = gcd(block widths) / groups (SD2.1: 10)
        import math
        self.stats_gran = math.gcd(*self.cfg["block_out_channels"]) // self.groups
        self._lora_token = None
        self.step_cache = {}          # CUDA-graph step states of the pipelines that share this UNet (pipeline.py)
        self._workspace = None
        self._pack(state_dict)
        self.set_lora(None)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_random(cls, seed: int = 0, config: dict = UNET_CONFIG, device="cuda:0"):
        return cls(random_state_dict(unet_manifest(config), seed), config, device)

    def _dev(self, t, dtype=f32):
        return t.to(device=self.device, dtype=dtype).contiguous()

    def _pack_resnet(self, sd, p: str, temb_rows: List[torch.Tensor], temb_bias: List[torch.Tensor]) -> _Resnet:
        r = _Resnet()
        w1 = sd[p + ".conv1.weight"]
        r.cout, r.cin = w1.shape[0], w1.shape[1]
        r.g1, r.b1 = self._dev(sd[p + ".norm1.weight"]), self._dev(sd[p + ".norm1.bias"])
        r.w1 = pack_conv_weight(w1, device=self.device)
        r.bias1 = self._dev(sd[p + ".conv1.bias"])
        r.temb_off = sum(t.shape[0] for t in temb_rows)
        temb_rows.append(sd[p + ".time_emb_proj.weight"])
        temb_bias.append(sd[p + ".time_emb_proj.bias"])
        r.g2, r.b2 = self._dev(sd[p + ".norm2.weight"]), self._dev(sd[p + ".norm2.bias"])
        sc = sd.get(p + ".conv_shortcut.weight")
        r.shortcut = sc is not None
        r.w2 = pack_conv_weight(sd[p + ".conv2.weight"], shortcut=sc, device=self.device)
        b2 = sd[p + ".conv2.bias"].float()
        if r.shortcut:
            b2 = b2 + sd[p + ".conv_shortcut.bias"].float()
        r.bias2 = self._dev(b2)
        return r

    def _pack_transformer(self, sd, p: str) -> _Transformer:
        t = _Transformer()
        t.path = p
        tb = p + ".transformer_blocks.0"
        t.c = sd[p + ".proj_in.weight"].shape[0]
        t.heads = t.c // HEAD_DIM
        t.gn_g, t.gn_b = self._dev(sd[p + ".norm.weight"]), self._dev(sd[p + ".norm.bias"])
        t.w_in, t.b_in = self._dev(sd[p + ".proj_in.weight"], bf16), self._dev(sd[p + ".proj_in.bias"])
        t.w_out, t.b_out = self._dev(sd[p + ".proj_out.weight"], bf16), self._dev(sd[p + ".proj_out.bias"])
        t.ln = [(self._dev(sd[f"{tb}.norm{i}.weight"]), self._dev(sd[f"{tb}.norm{i}.bias"])) for i in (1, 2, 3)]
        a1, a2 = tb + ".attn1", tb + ".attn2"
        t.w_qkv = self._dev(torch.cat([sd[a1 + ".to_q.weight"], sd[a1 + ".to_k.weight"], sd[a1 + ".to_v.weight"]], 0), bf16)
        t.w_o1, t.b_o1 = self._dev(sd[a1 + ".to_out.0.weight"], bf16), self._dev(sd[a1 + ".to_out.0.bias"])
        t.w_q2 = self._dev(sd[a2 + ".to_q.weight"], bf16)
        t.w_kv2 = self._dev(torch.cat([sd[a2 + ".to_k.weight"], sd[a2 + ".to_v.weight"]], 0), bf16)
        t.w_o2, t.b_o2 = self._dev(sd[a2 + ".to_out.0.weight"], bf16), self._dev(sd[a2 + ".to_out.0.bias"])
        wi, bi = interleave_geglu(sd[tb + ".ff.net.0.proj.weight"], sd[tb + ".ff.net.0.proj.bias"])
        t.w_ff1, t.b_ff1 = self._dev(wi, bf16), self._dev(bi)
        t.w_ff2, t.b_ff2 = self._dev(sd[tb + ".ff.net.2.weight"], bf16), self._dev(sd[tb + ".ff.net.2.bias"])
        t.lora = {}
        return t

    def _pack(self, sd):
        cfg = self.cfg
        ch = tuple(cfg["block_out_channels"])
        n = len(ch)
        temb_rows, temb_bias = [], []
        # conv_in (4 -> 320) on the tensor-core kernel: the 4 input channels ride in one 64-channel K block as
        # [x_hi | x_lo | x_hi] against [w_hi | w_hi | w_lo] (ops.latent_operand), i.e. at fp32-product precision
        wi = sd["conv_in.weight"].float().permute(0, 2, 3, 1)                     # [320, 3, 3, 4]
        w_hi = wi.to(bf16).float()
        wp = torch.zeros(wi.shape[0], 3, 3, 64)
        wp[..., 0:4], wp[..., 4:8], wp[..., 8:12] = w_hi, w_hi, wi - w_hi
        self.w_conv_in = wp.reshape(wi.shape[0], -1).to(device=self.device, dtype=bf16).contiguous()
        self.b_conv_in = self._dev(sd["conv_in.bias"])
        self.t_w1, self.t_b1 = self._dev(sd["time_embedding.linear_1.weight"]), self._dev(sd["time_embedding.linear_1.bias"])
        self.t_w2, self.t_b2 = self._dev(sd["time_embedding.linear_2.weight"]), self._dev(sd["time_embedding.linear_2.bias"])
        self.down, self.up = [], []
        for i in range(n):
            attn = cfg["down_block_types"][i].startswith("CrossAttn")
            blk = SimpleNamespace(resnets=[], attns=[], down=None)
            for j in range(cfg["layers_per_block"]):
                blk.resnets.append(self._pack_resnet(sd, f"down_blocks.{i}.resnets.{j}", temb_rows, temb_bias))
                if attn:
                    blk.attns.append(self._pack_transformer(sd, f"down_blocks.{i}.attentions.{j}"))
            if i < n - 1:
                q = f"down_blocks.{i}.downsamplers.0.conv"
                blk.down = (pack_conv_weight(sd[q + ".weight"], device=self.device), self._dev(sd[q + ".bias"]))
            self.down.append(blk)
        self.mid = SimpleNamespace(
            res0=self._pack_resnet(sd, "mid_block.resnets.0", temb_rows, temb_bias),
            attn=self._pack_transformer(sd, "mid_block.attentions.0"),
            res1=self._pack_resnet(sd, "mid_block.resnets.1", temb_rows, temb_bias))
        for i in range(n):
            attn = cfg["up_block_types"][i].startswith("CrossAttn")
            blk = SimpleNamespace(resnets=[], attns=[], up=None)
            for j in range(cfg["layers_per_block"] + 1):
                blk.resnets.append(self._pack_resnet(sd, f"up_blocks.{i}.resnets.{j}", temb_rows, temb_bias))
                if attn:
                    blk.attns.append(self._pack_transformer(sd, f"up_blocks.{i}.attentions.{j}"))
            if i < n - 1:
                q = f"up_blocks.{i}.upsamplers.0.conv"
                blk.up = (pack_conv_weight(sd[q + ".weight"], device=self.device), self._dev(sd[q + ".bias"]),
                          *reversed(pack_upsample_phase_weights(sd[q + ".weight"], device=self.device, stacked=True)))   # [2]: per-phase views, [3]: stacked
            self.up.append(blk)
        self.t_w_all = self._dev(torch.cat(temb_rows, 0))
        self.t_b_all = self._dev(torch.cat(temb_bias, 0))
        self.out_g, self.out_b = self._dev(sd["conv_norm_out.weight"]), self._dev(sd["conv_norm_out.bias"])
        # conv_out (320 -> 4) on the tensor-core kernel: Cout zero-padded to one 32-column epilogue chunk.  The spare
        # rows carry the bf16 rounding residual of the weights (w = hi + lo), so the last layer keeps ~16 mantissa bits
        # of weight precision for free: eps = out[:, 0:4] + out[:, 4:8]
        wo32 = sd["conv_out.weight"].float().permute(0, 2, 3, 1).reshape(sd["conv_out.weight"].shape[0], -1)
        hi = wo32.to(bf16)
        lo = (wo32 - hi.float()).to(bf16)
        co = wo32.shape[0]
        self.w_conv_out = torch.zeros((32, wo32.shape[1]), dtype=bf16, device=self.device)
        self.w_conv_out[:co] = hi.to(self.device)
        self.w_conv_out[co:2 * co] = lo.to(self.device)
        self.b_conv_out = torch.zeros(32, dtype=f32, device=self.device)
        self.b_conv_out[:co] = sd["conv_out.bias"].to(self.device)
        self.out_channels = co
        self.transformers: List[_Transformer] = [a for b in self.down for a in b.attns] + [self.mid.attn] + \
            [a for b in self.up for a in b.attns]
        self.cross_dim = cfg["cross_attention_dim"]

    # ------------------------------------------------------------------ LoRA (peft lora.Linear, unmerged)
    def set_lora(self, lora: Optional[dict], token=None) -> None:
        """lora: {module_path: (down [r,in], up [out,r], scale)} or None.  Only the small packed
        adapter tensors change; base weights stay bit-identical (adapter hot-swap).

        The packed adapter buffers are PERSISTENT: a second adapter set of the same layout is copied into the same
        device buffers, so CUDA graphs captured with adapters installed stay valid across `load_lora_weights`
        (`/root/reference/inference_ID-Booth.py:103-107` swaps adapters every 21 images).  `lora_topology` (which fused
        projections carry adapters) is what a captured graph depends on; `token` identifies the owner of the installed
        set (several pipelines may share this UNet)."""
        known = set()
        topo = []
        for t in self.transformers:
            tb = t.path + ".transformer_blocks.0"
            get = (lambda k: lora.get(k)) if lora else (lambda k: None)
            a1, a2 = tb + ".attn1", tb + ".attn2"
            known.update(f"{a}.{m}" for a in (a1, a2) for m in ("to_q", "to_k", "to_v", "to_out.0"))
            groups = {
                "qkv": ([get(a1 + ".to_q"), get(a1 + ".to_k"), get(a1 + ".to_v")], t.c),
                "o1": ([get(a1 + ".to_out.0")], t.c),
                "q2": ([get(a2 + ".to_q")], t.c),
                "kv2": ([get(a2 + ".to_k"), get(a2 + ".to_v")], self.cross_dim),
                "o2": ([get(a2 + ".to_out.0")], t.c),
            }
            bufs = t.__dict__.setdefault("lora_buf", {})
            active = {}
            for name, (ads, k) in groups.items():
                live = [a for a in ads if a is not None]
                if not live:
                    active[name] = (None, None)       # the buffers (if any) stay alive for graphs captured with them
                    topo.append(False)
                    continue
                old = bufs.get(name)
                if old is not None and all(a is not None and a[0].shape[0] <= 16 for a in ads):
                    # in place: two slice copies per adapter straight into the persistent packed buffers (graphs captured
                    # with them stay valid; a training loop re-installs 128 adapters per step this way)
                    ld, lu = old
                    for s_, (d, u, scale) in enumerate(ads):
                        r = d.shape[0]
                        ld[s_ * 16:s_ * 16 + r].copy_(d.detach())
                        if r < 16:
                            ld[s_ * 16 + r:(s_ + 1) * 16].zero_()
                        lu[s_ * t.c:(s_ + 1) * t.c, :r].copy_(u.detach() * float(scale))
                        if r < lu.shape[1]:
                            lu[s_ * t.c:(s_ + 1) * t.c, r:].zero_()
                else:
                    ld, lu = pack_lora(ads, self.device, seg_n=t.c, k=k)
                    if old is not None and (old[0].shape != ld.shape or old[1].shape != lu.shape):
                        self.step_cache.clear()   # captured graphs hold the old buffers' addresses
                    if old is not None and old[0].shape == ld.shape and old[1].shape == lu.shape:
                        old[0].copy_(ld)
                        old[1].copy_(lu)
                    else:
                        bufs[name] = (ld, lu)
                active[name] = bufs[name]
                topo.append(True)
            t.lora = active
        if lora:
            unknown = set(lora) - known
            if unknown:
                raise KeyError(f"LoRA targets not present in this UNet: {sorted(unknown)[:4]} ...")
        self.lora_topology = tuple(topo)
        self._lora_token = token
        self._lora_version += 1

    # ------------------------------------------------------------------ helpers
    def _ws(self):
        if self._workspace is None:
            self._workspace = torch.empty(WORKSPACE_BYTES // 4, dtype=f32, device=self.device)
        return self._workspace

    def _gemm(self, a0, w, **kw):
        return ops.gemm_conv(a0, w, k_splits=0, workspace=self._ws(), sums_pool=self._sums, stats_gran=self.stats_gran, **kw)

    def _lin_lora(self, a0, w, lo, seg_n, **kw):
        ld, lu = lo
        if ld is None:
            return self._gemm(a0, w, **kw)
        return self._gemm(a0, w, lora_down=ld, lora_up=lu, lora_seg_n=seg_n, **kw)

    # A "stream" value is (fp32 NHWC tensor, row-block channel statistics or None): every GEMM that
    # produces a residual-stream tensor also emits the sums GroupNorm needs, so the norm reads it once.
    def _resnet(self, r: _Resnet, hs, skip, temb, gnws):
        h, h_st = hs[0], hs[1]
        h_ph = hs[2] if len(hs) > 2 else 0          # 4: statistics written by the four phased upsample GEMMs
        sk, sk_st = skip if skip is not None else (None, None)
        B, H, W = h.shape[0], h.shape[1], h.shape[2]
        n1, raw = ops.groupnorm(h, r.g1, r.b1, groups=self.groups, eps=self.eps, silu=True, x1=sk,
                                want_raw=r.shortcut, partials=gnws, x0_stats=h_st, x1_stats=sk_st, x0_stats_phases=h_ph)
        t1, _, t1_st = self._gemm(n1, r.w1, mode=ops.A_3X3, bias=r.bias1, rowvec=temb[:, r.temb_off:],
                                  rowvec_ld=temb.shape[1], want_f32=True, want_stats=True)
        t1 = t1.view(B, H, W, r.cout)
        n2, _ = ops.groupnorm(t1, r.g2, r.b2, groups=self.groups, eps=self.eps, silu=True, partials=gnws, x0_stats=t1_st)
        if r.shortcut:
            o, _, o_st = self._gemm(n2, r.w2, mode=ops.A_3X3, a1=raw, bias=r.bias2, want_f32=True, want_stats=True)
        else:
            o, _, o_st = self._gemm(n2, r.w2, mode=ops.A_3X3, bias=r.bias2, residual=h, want_f32=True, want_stats=True)
        return o.view(B, H, W, r.cout), o_st

    def _context_kv(self, t: _Transformer, ctx_bf16):
        _, kv = self._lin_lora(ctx_bf16, t.w_kv2, t.lora["kv2"], t.c, want_bf16=True)
        return kv

    def _transformer(self, t: _Transformer, hs, kv, n_ctx, gnws):
        h, h_st = hs
        B, H, W, Cc = h.shape
        M, T = B * H * W, H * W
        n, _ = ops.groupnorm(h, t.gn_g, t.gn_b, groups=self.groups, eps=1e-6, silu=False, partials=gnws, x0_stats=h_st)
        x0, _ = self._gemm(n.view(M, Cc), t.w_in, bias=t.b_in, want_f32=True)
        # self-attention
        a = ops.layernorm(x0, *t.ln[0])
        _, qkv = self._lin_lora(a, t.w_qkv, t.lora["qkv"], Cc, want_bf16=True)
        o = ops.attention(qkv, qkv, qkv, batch=B, heads=t.heads, t_q=T, t_kv=T, scale=HEAD_DIM ** -0.5,
                          col0_q=0, col0_k=Cc, col0_v=2 * Cc)
        x1, _ = self._lin_lora(o, t.w_o1, t.lora["o1"], Cc, bias=t.b_o1, residual=x0, want_f32=True)
        # cross-attention over the (step-invariant) projected context
        a = ops.layernorm(x1, *t.ln[1])
        _, q = self._lin_lora(a, t.w_q2, t.lora["q2"], Cc, want_bf16=True)
        o = ops.attention(q, kv, kv, batch=B, heads=t.heads, t_q=T, t_kv=n_ctx, scale=HEAD_DIM ** -0.5,
                          col0_q=0, col0_k=0, col0_v=Cc)
        x2, _ = self._lin_lora(o, t.w_o2, t.lora["o2"], Cc, bias=t.b_o2, residual=x1, want_f32=True)
        # GEGLU feed-forward
        a = ops.layernorm(x2, *t.ln[2])
        _, g = self._gemm(a, t.w_ff1, bias=t.b_ff1, geglu=True, want_bf16=True)
        _, x3 = self._gemm(g, t.w_ff2, bias=t.b_ff2, residual=x2, want_bf16=True)
        out, _, out_st = self._gemm(x3, t.w_out, bias=t.b_out, residual=h.view(M, Cc), want_f32=True, want_stats=True,
                                    stats_hw=T)
        return out.view(B, H, W, Cc), out_st

    def _transformer_shared_prefix(self, t: _Transformer, hs, kv, n_ctx, gnws):
        """The FIRST transformer of a classifier-free-guidance forward: both halves of the CFG pair carry the same latent,
        so everything up to the cross-attention (GroupNorm, proj_in, self-attention, to_out, the cross-attention queries)
        is computed ONCE on the n shared images; the two text contexts split the stream at the cross-attention, and from
        there on the batch is 2n.  hs: the n-image stream; kv: projected context of all 2n rows.  Bit-identical to running
        the block on the duplicated batch (every kernel on the shared part is per-image / per-row)."""
        h, h_st = hs
        n, H, W, Cc = h.shape
        M, T = n * H * W, H * W
        g, _ = ops.groupnorm(h, t.gn_g, t.gn_b, groups=self.groups, eps=1e-6, silu=False, partials=gnws, x0_stats=h_st)
        x0, _ = self._gemm(g.view(M, Cc), t.w_in, bias=t.b_in, want_f32=True)
        a = ops.layernorm(x0, *t.ln[0])
        _, qkv = self._lin_lora(a, t.w_qkv, t.lora["qkv"], Cc, want_bf16=True)
        o = ops.attention(qkv, qkv, qkv, batch=n, heads=t.heads, t_q=T, t_kv=T, scale=HEAD_DIM ** -0.5,
                          col0_q=0, col0_k=Cc, col0_v=2 * Cc)
        x1, _ = self._lin_lora(o, t.w_o1, t.lora["o1"], Cc, bias=t.b_o1, residual=x0, want_f32=True)
        a = ops.layernorm(x1, *t.ln[1])
        _, q = self._lin_lora(a, t.w_q2, t.lora["q2"], Cc, want_bf16=True)
        # ---- the stream splits here: same queries, two contexts; the residuals x1 / h are read by both halves
        o2 = torch.empty((2 * M, Cc), dtype=bf16, device=self.device)
        x2 = torch.empty((2 * M, Cc), dtype=f32, device=self.device)
        for half in range(2):
            kvh = kv[half * n * n_ctx:(half + 1) * n * n_ctx]
            ops.attention(q, kvh, kvh, o2[half * M:(half + 1) * M], batch=n, heads=t.heads, t_q=T, t_kv=n_ctx,
                          scale=HEAD_DIM ** -0.5, col0_q=0, col0_k=0, col0_v=Cc)
            self._lin_lora(o2[half * M:(half + 1) * M], t.w_o2, t.lora["o2"], Cc, bias=t.b_o2, residual=x1,
                           out_f32=x2[half * M:(half + 1) * M])
        a = ops.layernorm(x2, *t.ln[2])
        _, gg = self._gemm(a, t.w_ff1, bias=t.b_ff1, geglu=True, want_bf16=True)
        _, x3 = self._gemm(gg, t.w_ff2, bias=t.b_ff2, residual=x2, want_bf16=True)
        out = torch.empty((2 * M, Cc), dtype=f32, device=self.device)
        use_sums = ops.epilogue_stats_supported(1, 1, 2 * M) and ops.image_sums_supported(2 * n, T) and Cc % self.stats_gran == 0
        out_st = self._sums.take(2 * n, Cc // self.stats_gran) if use_sums else None
        for half in range(2):
            self._gemm(x3[half * M:(half + 1) * M], t.w_out, bias=t.b_out, residual=h.view(M, Cc), out_f32=out[half * M:(half + 1) * M],
                       sums=None if out_st is None else out_st[half * n:(half + 1) * n], stats_hw=T)
        return out.view(2 * n, H, W, Cc), out_st

    # ------------------------------------------------------------------ step-invariant context projections
    def encode_context(self, encoder_hidden_states: torch.Tensor):
        """Cross-attention K/V projections (incl. their LoRA deltas) of the text context: they do
        not depend on the timestep, so the pipeline computes them once per prompt batch."""
        B, S, D = encoder_hidden_states.shape
        ctx = encoder_hidden_states.to(device=self.device, dtype=bf16).reshape(B * S, D).contiguous()
        return SimpleNamespace(kv=[self._context_kv(t, ctx) for t in self.transformers], n_ctx=S, batch=B,
                               lora_version=self._lora_version)

    # ------------------------------------------------------------------ timestep-only part of the forward
    def time_embedding(self, timesteps: torch.Tensor) -> torch.Tensor:
        """Timesteps fp32 [S] -> [S, sum of ResnetBlock2D widths]: sinusoid -> TimestepEmbedding MLP -> SiLU ->
        all 22 `time_emb_proj` layers in one batched call.  It depends on the timestep only (not on the
        sample, the prompt or LoRA), so the pipeline evaluates it once for all scheduler timesteps of a call
        and hands each step its row (`forward(..., temb=...)`)."""
        t = timesteps.to(device=self.device, dtype=f32).reshape(-1).contiguous()
        return ops.time_embed(t, self.t_w1, self.t_b1, self.t_w2, self.t_b2, self.t_w_all, self.t_b_all)

    # ------------------------------------------------------------------ forward
    def forward(self, sample, timestep, encoder_hidden_states=None, class_labels=None, return_dict: bool = True,
                context=None, taps: Optional[dict] = None, temb: Optional[torch.Tensor] = None, cfg_pair: bool = False):
        """`cfg_pair=True` (an extension the pipeline uses): `sample` holds the n latents ONCE while the context holds 2n
        rows ([uncond | cond], `pipeline_stable_diffusion.py` feeds `torch.cat([latents] * 2)`); the result has 2n rows
        and equals `forward(torch.cat([sample] * 2), ...)` bit for bit -- conv_in, the first ResnetBlock2D and the first
        transformer up to its cross-attention see identical inputs in both halves and are evaluated once."""
        if class_labels is not None:
            raise NotImplementedError("SD2.1-base has no class embedding")
        in_dtype = sample.dtype
        x = sample.to(device=self.device, dtype=f32).contiguous()
        n_shared = x.shape[0] if cfg_pair else 0
        if cfg_pair and not (self.down and self.down[0].attns):
            raise NotImplementedError("cfg_pair needs a cross-attention block right after conv_in")
        B, Cin, H, W = x.shape
        if cfg_pair:
            B = 2 * n_shared            # batch of the result (and of everything after the first cross-attention)
        if H % 8 or W % 8:
            raise ValueError("latent height/width must be multiples of 8 (three stride-2 stages)")
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([float(timestep)], dtype=f32, device=self.device)
        t = timestep.to(device=self.device, dtype=f32).reshape(-1)
        if t.numel() == 1:
            t = t.expand(B)
        elif cfg_pair and t.numel() == n_shared:
            t = t.repeat(2)
        t = t.contiguous()
        if context is None:
            context = self.encode_context(encoder_hidden_states)
        if context.lora_version != self._lora_version or context.batch != B:
            raise RuntimeError("stale / mismatched encode_context() result")
        kvs = iter(context.kv)
        gnws = ops.groupnorm_workspace(B, self.groups, self.device)
        # per-image GroupNorm sums of every residual-stream tensor of this forward: one zeroed pool (one memset), a slice per GEMM
        self._sums = ops.SumsPool(self.device, capacity=B * 2 * 61440)
        if temb is None:
            temb = ops.time_embed(t, self.t_w1, self.t_b1, self.t_w2, self.t_b2, self.t_w_all, self.t_b_all)
        elif temb.shape != (B, self.t_w_all.shape[0]) or temb.dtype != f32 or not temb.is_contiguous():
            raise ValueError("temb must be a contiguous fp32 [batch, n_time_proj] tensor from time_embedding()")

        def tap(name, v):
            if taps is not None:
                taps[name] = v[0].permute(0, 3, 1, 2).float().clone()

        h0, _, h0_st = self._gemm(ops.latent_operand(x), self.w_conv_in, mode=ops.A_3X3, bias=self.b_conv_in, want_f32=True,
                                  want_stats=True)
        h = (h0.view(x.shape[0], H, W, -1), h0_st)
        tap("conv_in", h)
        skips = [h]                      # (cfg_pair: this skip holds n images; its consumer reads image b % n)
        for i, blk in enumerate(self.down):
            for j, r in enumerate(blk.resnets):
                if cfg_pair and i == 0 and j == 0:
                    h = self._resnet(r, h, None, temb[:n_shared], gnws)     # (both halves share the timestep rows)
                    h = self._transformer_shared_prefix(blk.attns[0], h, next(kvs), context.n_ctx, gnws)
                    skips.append(h)
                    continue
                h = self._resnet(r, h, None, temb, gnws)
                tap(f"down_blocks.{i}.resnets.{j}", h)
                if blk.attns:
                    h = self._transformer(blk.attns[j], h, next(kvs), context.n_ctx, gnws)
                    tap(f"down_blocks.{i}.attentions.{j}", h)
                skips.append(h)
            if blk.down is not None:
                ht = h[0]
                hb = ops.cast_bf16(ht)
                o, _, o_st = self._gemm(hb, blk.down[0], mode=ops.A_3X3_S2, bias=blk.down[1], want_f32=True, want_stats=True)
                h = (o.view(B, ht.shape[1] // 2, ht.shape[2] // 2, ht.shape[3]), o_st)
                skips.append(h)
        h = self._resnet(self.mid.res0, h, None, temb, gnws)
        h = self._transformer(self.mid.attn, h, next(kvs), context.n_ctx, gnws)
        h = self._resnet(self.mid.res1, h, None, temb, gnws)
        tap("mid_block", h)
        for i, blk in enumerate(self.up):
            for j, r in enumerate(blk.resnets):
                h = self._resnet(r, h, skips.pop(), temb, gnws)
                if blk.attns:
                    h = self._transformer(blk.attns[j], h, next(kvs), context.n_ctx, gnws)
                tap(f"up_blocks.{i}.{j}", h)
            if blk.up is not None:
                ht = h[0]
                Hl, Wl, Cu = ht.shape[1], ht.shape[2], ht.shape[3]
                if B * Hl * Wl >= 2048 and ops.epilogue_stats_supported(B, Hl, Wl):
                    # Upsample2D as four 2x2 convolutions on the LOW-resolution tensor (one per output parity class,
                    # 3x3 taps pre-summed): 4 x K = 4C instead of K = 9C on the upsampled tensor -> 2.25x fewer MACs,
                    # and no upsampled operand is ever written
                    xb = ops.cast_bf16(ht)
                    o = torch.empty((B, 2 * Hl, 2 * Wl, Cu), dtype=f32, device=self.device)
                    o_st = torch.empty((4, B * Hl * Wl // 32, Cu, 2), dtype=f32, device=self.device)
                    # per-image channel sums accumulated by the four phase calls (else: phased row-block sums)
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
        n, _ = ops.groupnorm(h[0], self.out_g, self.out_b, groups=self.groups, eps=self.eps, silu=True, partials=gnws,
                             x0_stats=h[1])
        o, _ = self._gemm(n, self.w_conv_out, mode=ops.A_3X3, bias=self.b_conv_out, want_f32=True)
        o = o.view(B, H, W, 32)
        co = self.out_channels
        eps = (o[..., :co] + o[..., co:2 * co]).permute(0, 3, 1, 2).contiguous()   # hi + lo weight halves; NHWC -> NCHW
        if in_dtype != f32:
            eps = eps.to(in_dtype)
        return UNetOutput(eps) if return_dict else (eps,)

    __call__ = forward

    # diffusers-compat no-ops
    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def train(self, mode: bool = True):
        return self  # dropout p = 0: train-mode forward is the same arithmetic (train_ID-Booth.py:989)

    def requires_grad_(self, flag: bool = False):
        return self
