"""Tensor-level wrappers over the C ABI (include/idb.h).  Each wrapper only validates
dtype / layout / device, fills the POD argument struct with raw device pointers and
launches on torch's current CUDA stream.  Outputs are caller-provided (or allocated
with torch when omitted) -- the library itself never allocates.
"""
from __future__ import annotations

import ctypes as C
import ctypes as C_
import os
from typing import Optional

import torch

from . import _lib
from ._lib import A_1X1, A_2X2, A_3X3, A_3X3_S2, A_3X3_S2_ASYM, EPI_F16, EPI_GEGLU, EPI_GELU, EPI_PHASES4  # noqa: F401  (re-exported)

bf16, f16, f32 = torch.bfloat16, torch.float16, torch.float32


def _chk(t: Optional[torch.Tensor], dtype, name: str, allow_none: bool = False):
    if t is None:
        if allow_none:
            return
        raise ValueError(f"{name} is required")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.device.index != torch.cuda.current_device():
        # launches go to the CURRENT device's current stream: a tensor of another GPU would be read by the wrong device
        raise RuntimeError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                           "wrap the call in `with torch.cuda.device(tensor.device):`")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


_IMAGE_SUMS_ON = os.environ.get("IDB_IMAGE_SUMS", "1") != "0"   # 0: row-block sums + finalize kernel (A/B profiling)
PHASES4_ON = os.environ.get("IDB_PHASES4", "1") != "0"   # Upsample2D: the four phase GEMMs as one launch (0: four launches)
IMAGE_SUMS_MAX_ROWS = 16384     # rows per image up to which the fixed-point per-image sums are used (range: see idb.h)


class SumsPool:
    """Zero-initialised int64 storage for the per-image channel sums (`stats_image_sums`) of MANY GEMMs: the GEMM epilogues
    ADD into their slice with integer atomics, so every slice must start at zero -- one memset per forward instead of one
    per GEMM.  A pool is used once (its slices are consumed by the GroupNorms of the same forward)."""

    def __init__(self, device, capacity: int = 1 << 20):
        self.device, self.capacity = device, capacity
        self.buf, self.used = None, 0

    def take(self, n_images: int, n: int) -> torch.Tensor:
        """int64 [n_images, n, 2], zero (n = number of channel granules)"""
        need = n_images * n * 2
        if self.buf is None or self.used + need > self.buf.numel():
            self.buf = torch.zeros(max(self.capacity, need), dtype=torch.int64, device=self.device)
            self.used = 0
        out = self.buf[self.used:self.used + need].view(n_images, n, 2)
        self.used += need
        return out


def image_sums_supported(n_images: int, rows_per_image: int, n: int = 0, phased: bool = False) -> bool:
    """Mirror of the `stats_image_sums` precondition of `idb_gemm_conv` that does not depend on the tile geometry: whole
    32-row blocks per image, and few enough rows that the fixed-point accumulators cannot overflow on sane activations
    (`rows_per_image` is per phase when the output is written by four phased calls)."""
    rows = 4 * rows_per_image if phased else rows_per_image
    return _IMAGE_SUMS_ON and rows_per_image % 32 == 0 and rows <= IMAGE_SUMS_MAX_ROWS


def gemm_conv(a0: torch.Tensor, w: torch.Tensor, *, mode: int = A_1X1, a1: Optional[torch.Tensor] = None,
              bias=None, rowvec=None, rowvec_ld: int = 0, residual=None, lora_down=None, lora_up=None, lora_seg_n: int = 0,
              geglu: bool = False, out_f32: Optional[torch.Tensor] = None, out_bf16: Optional[torch.Tensor] = None,
              want_f32: bool = False, want_bf16: bool = False, k_splits: int = 1,
              workspace: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None, want_stats: bool = False,
              prelu: Optional[torch.Tensor] = None, half: bool = False, gelu: bool = False,
              tap_off=(0, 0), out_phase=None, sums: Optional[torch.Tensor] = None, stats_hw: int = 0,
              sums_pool: Optional[SumsPool] = None, stats_gran: int = 1, phases4: bool = False):
    """a0: bf16 [B,H,W,C0] (or [M,K] for a Linear); w: bf16 [N, Ktot].  Returns (out_f32, out_bf16).
    half=True: the 16-bit tensors (a0, a1, w, out_bf16) are IEEE fp16 instead of bf16 (IDB_EPI_F16).

    want_stats=True additionally returns the GroupNorm statistics of the fp32 output, in the best form the geometry allows:
    an int64 [n_images, N / stats_gran, 2] tensor of fixed-point per-image sums over granules of `stats_gran` channels
    (`stats_image_sums`, accumulated by integer atomics in the epilogue; taken from `sums_pool` or zero-allocated;
    `stats_hw` = rows per image when the operand is a token matrix; `stats_gran` must divide the group size of every
    GroupNorm consuming the tensor), else a float32 row-block tensor (`stats_partials`), else None (GroupNorm then computes its own).
    `groupnorm(x0_stats=...)` accepts any of the three.  An explicit `sums=` must be zero before the (first phase) call."""
    t16 = f16 if half else bf16
    _chk(a0, t16, "a0"); _chk(w, t16, "w")
    if a0.dim() == 2:
        B, H, W_, C0 = 1, 1, a0.shape[0], a0.shape[1]
    else:
        B, H, W_, C0 = a0.shape
    N = w.shape[0]
    Ho, Wo = (H // 2, W_ // 2) if mode in (A_3X3_S2, A_3X3_S2_ASYM) else (H, W_)
    M = B * Ho * Wo
    taps = 1 if mode == A_1X1 else (4 if mode == A_2X2 else 9)
    c1 = 0
    if a1 is not None:
        _chk(a1, t16, "a1")
        c1 = a1.shape[-1]
        if a1.numel() != M * c1:
            raise ValueError("a1 must have the output geometry")
    if w.shape[1] != taps * C0 + c1:
        raise ValueError(f"w has K={w.shape[1]}, expected {taps * C0 + c1}")
    # phases4 (A_2X2): w holds the four phase matrices of an Upsample2D stacked on N; one call writes all four output parity
    # classes of out_f32 [B, 2H, 2W, N / 4] (IDB_EPI_PHASES4); `stats` is then [4, M / 32, N / 4, 2], `sums` as usual
    if phases4 and (mode != A_2X2 or out_phase is not None or N % 4 or (out_f32 is None and out_bf16 is None)):
        raise ValueError("phases4 needs mode A_2X2, no out_phase, N = 4 * N_out and an explicit output tensor")
    n_out = N // 2 if geglu else (N // 4 if phases4 else N)
    for t, nm in ((bias, "bias"), (residual, "residual"), (prelu, "prelu")):
        _chk(t, f32, nm, allow_none=True)
    if prelu is not None and prelu.numel() != N:
        raise ValueError("prelu must hold one slope per output channel")
    _chk(lora_up, bf16, "lora_up", allow_none=True)
    if lora_up is not None and tuple(lora_up.shape) != (N, 64):
        raise ValueError("lora_up must be bf16 [N, 64] (packing.pack_lora)")
    if rowvec is not None:  # may be a column slice of a wider [batch, ld] table (rowvec_ld = ld)
        if not rowvec.is_cuda or rowvec.dtype != f32 or rowvec.stride(-1) != 1:
            raise ValueError("rowvec must be a CUDA fp32 tensor with unit inner stride")
        if rowvec_ld and rowvec.dim() == 2 and rowvec.stride(0) != rowvec_ld:
            raise ValueError("rowvec_ld does not match the row stride of rowvec")
    _chk(lora_down, bf16, "lora_down", allow_none=True)
    if residual is not None and residual.numel() != M * n_out:
        raise ValueError("residual must be [M, N_out]")
    if rowvec is not None and rowvec_ld == 0 and rowvec.numel() != B * N:
        raise ValueError("rowvec must be [batch, N]")
    if out_f32 is None and want_f32:
        out_f32 = torch.empty((M, n_out), dtype=f32, device=a0.device)
    if out_bf16 is None and (want_bf16 or (out_f32 is None)):
        out_bf16 = torch.empty((M, n_out), dtype=t16, device=a0.device)
    _chk(out_f32, f32, "out_f32", allow_none=True); _chk(out_bf16, t16, "out_bf16", allow_none=True)
    if k_splits > 1 and workspace is None:
        workspace = torch.empty((k_splits, M, N), dtype=f32, device=a0.device)
    # row-block channel statistics of the fp32 output (consumed by groupnorm); geometries whose 128-pixel tiles are not
    # raster runs of one image (e.g. the 96 / 48 / 24 / 12-wide rasters of 768 x 768 images) return stats = None and the
    # GroupNorm that follows computes its own statistics (`gn_stats_kernel`)
    if stats is None and sums is None and want_stats and epilogue_stats_supported(B, Ho, Wo):
        hw_img = stats_hw or Ho * Wo
        if out_phase is None and M % hw_img == 0 and n_out % stats_gran == 0 and image_sums_supported(M // hw_img, hw_img):
            sums = sums_pool.take(M // hw_img, n_out // stats_gran) if sums_pool is not None else \
                torch.zeros((M // hw_img, n_out // stats_gran, 2), dtype=torch.int64, device=a0.device)
        else:
            stats = torch.empty(((M + 31) // 32, n_out, 2), dtype=f32, device=a0.device)
    _chk(stats, f32, "stats", allow_none=True)
    _chk(sums, torch.int64, "sums", allow_none=True)
    args = _lib.GemmConvArgs(
        a0=a0.data_ptr(), a0_mode=mode, c0=C0, a1=_lib.ptr(a1), c1=c1, batch=B, height=H, width=W_,
        w=w.data_ptr(), n=N, bias=_lib.ptr(bias), rowvec=_lib.ptr(rowvec), rowvec_ld=rowvec_ld, residual=_lib.ptr(residual),
        lora_down=_lib.ptr(lora_down), lora_up=_lib.ptr(lora_up),
        lora_rank_pad=0 if lora_up is None else 16, lora_seg_n=lora_seg_n,
        flags=(EPI_GEGLU if geglu else 0) | (EPI_F16 if half else 0) | (EPI_GELU if gelu else 0) | (EPI_PHASES4 if phases4 else 0),
        out_f32=_lib.ptr(out_f32), out_bf16=_lib.ptr(out_bf16),
        k_splits=k_splits, workspace=_lib.ptr(workspace),
        workspace_bytes=0 if workspace is None else workspace.numel() * 4, stats_partials=_lib.ptr(stats),
        stats_image_sums=_lib.ptr(sums), stats_hw=stats_hw, stats_gran=stats_gran,
        prelu=_lib.ptr(prelu), tap_off_x=tap_off[1], tap_off_y=tap_off[0],
        out_scale=2 if (out_phase is not None or phases4) else 0, out_phase_y=0 if out_phase is None else out_phase[0],
        out_phase_x=0 if out_phase is None else out_phase[1])
    _lib.call("idb_gemm_conv", C.byref(args), _lib.stream_ptr(),
              desc=None if _lib.trace is None else dict(M=M, N=N, K=taps * C0 + c1, mode=mode, lora=lora_down is not None,
                                                        geglu=geglu, f32=out_f32 is not None, b16=out_bf16 is not None,
                                                        res=residual is not None, stats=stats is not None))
    if want_stats or stats is not None or sums is not None:
        return out_f32, out_bf16, (sums if sums is not None else stats)
    return out_f32, out_bf16


def _pow2_divisor(v: int, cap: int) -> int:
    d = 1
    while d * 2 <= cap and v % (d * 2) == 0:
        d *= 2
    return d


def epilogue_stats_supported(batch: int, ho: int, wo: int) -> bool:
    """Mirror of the `stats_partials` precondition of `idb_gemm_conv` (gemm_tc.cu): the 32-row blocks of a 128-pixel output
    tile (BW x BH x BB pixels) must be raster-contiguous runs of one image, and an image must hold whole blocks.
    A Linear ([M, K] operand) arrives here as batch = 1, ho = 1, wo = M."""
    if (ho * wo) % 32:
        return False
    if ho == 1 and batch == 1:
        return True
    bw = _pow2_divisor(wo, 128)
    bh = _pow2_divisor(ho, 128 // bw)
    bb = 128 // (bw * bh)
    return bw == wo or (bh == 1 and bb == 1)


def attention(q, k, v, out=None, *, batch: int, heads: int, t_q: int, t_kv: int, scale: float,
              col0_q: int = 0, col0_k: int = 0, col0_v: int = 0, causal: bool = False, lse: Optional[torch.Tensor] = None):
    """q: bf16 [B*Tq, ld_q]; k, v: bf16 [B*Tkv, ld]; head h lives at columns col0 + h*64.
    lse: optional fp32 [B, heads, Tq] output (log2-domain log-sum-exp of the scaled scores, for `attention_backward`)."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, bf16, nm)
    if out is None:
        out = torch.empty((batch * t_q, heads * 64), dtype=bf16, device=q.device)
    _chk(out, bf16, "out")
    _chk(lse, f32, "lse", allow_none=True)
    if lse is not None and lse.numel() != batch * heads * t_q:
        raise ValueError("lse must be fp32 [batch, heads, t_q]")
    args = _lib.AttentionArgs(q=q.data_ptr(), ld_q=q.shape[-1], col0_q=col0_q, k=k.data_ptr(), ld_k=k.shape[-1],
                              col0_k=col0_k, v=v.data_ptr(), ld_v=v.shape[-1], col0_v=col0_v, out=out.data_ptr(),
                              ld_out=out.shape[-1], batch=batch, heads=heads, t_q=t_q, t_kv=t_kv, scale=scale, causal=int(causal),
                              lse=_lib.ptr(lse))
    _lib.call("idb_attention", C.byref(args), _lib.stream_ptr(),
              desc=None if _lib.trace is None else dict(B=batch, heads=heads, Tq=t_q, Tkv=t_kv))
    return out


def attention_backward(q, k, v, o, d_o, lse, *, batch: int, heads: int, t_q: int, t_kv: int, scale: float,
                       col0_q: int = 0, col0_k: int = 0, col0_v: int = 0, dq=None, dk=None, dv=None, col0_dq: int = 0,
                       col0_dk: int = 0, col0_dv: int = 0):
    """Gradients of `attention` (same operand addressing; o / d_o are bf16 [B*Tq, heads*64]).  Returns (dq fp32
    [B*Tq, ...] -- accumulated atomically, so a caller-provided dq must be zero --, dk, dv bf16 [B*Tkv, ...])."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v"), (o, "o"), (d_o, "d_o")):
        _chk(t, bf16, nm)
    _chk(lse, f32, "lse")
    C = heads * 64
    if dq is None:
        dq = torch.zeros((batch * t_q, C), dtype=f32, device=q.device)
    if dk is None:
        dk = torch.empty((batch * t_kv, C), dtype=bf16, device=q.device)
    if dv is None:
        dv = torch.empty((batch * t_kv, C), dtype=bf16, device=q.device)
    _chk(dq, f32, "dq"); _chk(dk, bf16, "dk"); _chk(dv, bf16, "dv")
    dsum = torch.empty((batch, heads, t_q), dtype=f32, device=q.device)
    args = _lib.AttentionBwdArgs(
        q=q.data_ptr(), ld_q=q.shape[-1], col0_q=col0_q, k=k.data_ptr(), ld_k=k.shape[-1], col0_k=col0_k,
        v=v.data_ptr(), ld_v=v.shape[-1], col0_v=col0_v, o=o.data_ptr(), ld_o=o.shape[-1], col0_o=0,
        d_o=d_o.data_ptr(), ld_do=d_o.shape[-1], col0_do=0, lse=lse.data_ptr(), dsum=dsum.data_ptr(),
        dq=dq.data_ptr(), ld_dq=dq.shape[-1], col0_dq=col0_dq, dk=dk.data_ptr(), ld_dk=dk.shape[-1], col0_dk=col0_dk,
        dv=dv.data_ptr(), ld_dv=dv.shape[-1], col0_dv=col0_dv, batch=batch, heads=heads, t_q=t_q, t_kv=t_kv, scale=scale)
    _lib.call("idb_attention_backward", C_.byref(args), _lib.stream_ptr())
    return dq, dk, dv


def layernorm_backward(dy, x, gamma, dx=None, add: bool = False, eps: float = 1e-5):
    """dx (+)= d LayerNorm(x) / dx applied to dy; dy, x fp32 [rows, C]."""
    _chk(dy, f32, "dy"); _chk(x, f32, "x"); _chk(gamma, f32, "gamma")
    c = x.shape[-1]
    if dx is None:
        if add:
            raise ValueError("add=True needs dx")
        dx = torch.empty_like(x)
    _chk(dx, f32, "dx")
    _lib.call("idb_layernorm_backward", dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), dx.data_ptr(), int(add), x.numel() // c, c, eps,
              _lib.stream_ptr())
    return dx


def groupnorm_backward(dy, x0, gamma, beta, stats, *, groups: int, silu: bool, x1=None, dx0=None, dx1=None,
                       add0: bool = False, add1: bool = False, want_dx1: bool = True):
    """Input gradient of `groupnorm` over [x0 | x1]; dy fp32 [B, hw.., C0+C1]; stats fp32 [B, groups, 2] = (mean, rstd)."""
    _chk(dy, f32, "dy"); _chk(x0, f32, "x0"); _chk(x1, f32, "x1", allow_none=True); _chk(stats, f32, "stats")
    _chk(gamma, f32, "gamma"); _chk(beta, f32, "beta")
    B, c0 = x0.shape[0], x0.shape[-1]
    hw = x0.numel() // (B * c0)
    c1 = 0 if x1 is None else x1.shape[-1]
    if dx0 is None:
        dx0 = torch.empty_like(x0)
        add0 = False
    if x1 is not None and dx1 is None and want_dx1:
        dx1 = torch.empty_like(x1)
        add1 = False
    scratch = torch.empty((B, groups, 16, 2), dtype=f32, device=x0.device)     # up to 16 pixel slabs per (image, group)
    args = _lib.GroupNormBwdArgs(dy=dy.data_ptr(), x0=x0.data_ptr(), c0=c0, x1=_lib.ptr(x1), c1=c1, batch=B, hw=hw, groups=groups,
                                 silu=int(silu), stats=stats.data_ptr(), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                                 scratch=scratch.data_ptr(), dx0=_lib.ptr(dx0), dx1=_lib.ptr(dx1), add0=int(add0), add1=int(add1))
    _lib.call("idb_groupnorm_backward", C_.byref(args), _lib.stream_ptr())
    return dx0, dx1


def group_stats_from_sums(sums: torch.Tensor, channels: int, hw: int, groups: int, eps: float, sums1: Optional[torch.Tensor] = None,
                          channels1: int = 0) -> torch.Tensor:
    """(mean, rstd) fp32 [B, groups, 2] from the fixed-point per-image granule sums the forward GEMMs accumulated
    (`gemm_conv(want_stats=True)`), for `groupnorm_backward` -- the same arithmetic as the forward's apply kernel."""
    def dec(t, c):
        gran = c // t.shape[1]
        return t.double(), gran
    a, g0 = dec(sums, channels)
    parts = [a]
    gran = g0
    if sums1 is not None:
        b_, g1 = dec(sums1, channels1)
        if g1 != g0:
            raise ValueError("both sources must use the same channel granule")
        parts.append(b_)
    allg = torch.cat(parts, dim=1)                                    # [B, (C0 + C1) / gran, 2]
    C = channels + channels1
    per = (C // groups) // gran
    grp = allg.view(allg.shape[0], groups, per, 2).sum(2)
    n = float(hw) * (C // groups)
    mean = grp[..., 0] / 2.0 ** 32 / n
    var = (grp[..., 1] / 2.0 ** 24 / n - mean * mean).clamp_min(0.0)
    rstd = torch.rsqrt(var.float() + eps)
    return torch.stack([mean.float(), rstd], -1).contiguous()


def geglu_backward(dh, u, du=None):
    """dh bf16 [M, H]; u bf16 [M, 2H] interleaved pre-activation -> du bf16 [M, 2H]."""
    _chk(dh, bf16, "dh"); _chk(u, bf16, "u")
    M, H = dh.shape
    if du is None:
        du = torch.empty_like(u)
    _lib.call("idb_geglu_backward", dh.data_ptr(), u.data_ptr(), du.data_ptr(), M, H, _lib.stream_ptr())
    return du


_wgrad_ws = {}


def lora_wgrad(wide, skinny, rank: int, *, out=None, transpose_out: bool = False, scale: float = 1.0, add: bool = False,
               col0_w: int = 0, width: Optional[int] = None, col0_s: int = 0):
    """out[w, r] (or out[r, w]) (+)= scale * sum_m wide[m, col0_w + w] * skinny[m, col0_s + r]; wide / skinny bf16 2-D."""
    _chk(wide, bf16, "wide"); _chk(skinny, bf16, "skinny")
    M = wide.shape[0]
    W = width if width is not None else wide.shape[1] - col0_w
    if out is None:
        out = torch.zeros((rank, W) if transpose_out else (W, rank), dtype=f32, device=wide.device)
    _chk(out, f32, "out")
    key = (wide.device, W)
    if key not in _wgrad_ws:
        _wgrad_ws[key] = torch.empty(_lib.load().idb_lora_wgrad_workspace_bytes(W) // 4, dtype=f32, device=wide.device)
    _lib.call("idb_lora_wgrad", wide.data_ptr(), wide.shape[1], col0_w, skinny.data_ptr(), skinny.shape[1], col0_s, out.data_ptr(),
              out.shape[1], int(transpose_out), scale, int(add), M, W, rank, _wgrad_ws[key].data_ptr(), _lib.stream_ptr())
    return out


def zero_insert2x(g, out=None):
    _chk(g, bf16, "g")
    B, H, W_, c = g.shape
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W_, c), dtype=bf16, device=g.device)
    _lib.call("idb_zero_insert2x", g.data_ptr(), out.data_ptr(), B, H, W_, c, _lib.stream_ptr())
    return out


def sumpool2x(g, out=None, add: bool = False):
    _chk(g, f32, "g")
    B, H2, W2, c = g.shape
    if out is None:
        out = torch.empty((B, H2 // 2, W2 // 2, c), dtype=f32, device=g.device)
        add = False
    _lib.call("idb_sumpool2x", g.data_ptr(), out.data_ptr(), int(add), B, H2 // 2, W2 // 2, c, _lib.stream_ptr())
    return out


def groupnorm_workspace(batch: int, groups: int, device) -> torch.Tensor:
    n = _lib.load().idb_groupnorm_workspace_bytes(batch, groups)
    return torch.zeros(n // 4, dtype=f32, device=device)   # arrival counters must start at zero


def groupnorm(x0, gamma, beta, *, groups: int, eps: float, silu: bool, x1=None, out_norm=None, out_raw=None,
              want_raw: bool = False, partials=None, x0_stats=None, x1_stats=None, x0_stats_phases: int = 0):
    """x0: fp32 [B, HW.., C0] NHWC (+ optional x1 [B, HW.., C1] concatenated on channels)."""
    _chk(x0, f32, "x0"); _chk(x1, f32, "x1", allow_none=True)
    _chk(gamma, f32, "gamma"); _chk(beta, f32, "beta")
    B, c0 = x0.shape[0], x0.shape[-1]
    hw = x0.numel() // (B * c0)
    c1 = 0 if x1 is None else x1.shape[-1]
    Ct = c0 + c1
    shape = tuple(x0.shape[:-1]) + (Ct,)
    if out_norm is None:
        out_norm = torch.empty(shape, dtype=bf16, device=x0.device)
    if out_raw is None and want_raw:
        out_raw = torch.empty(shape, dtype=bf16, device=x0.device)
    if partials is None:
        partials = groupnorm_workspace(B, groups, x0.device)
    # statistics handed over by the producers: int64 = fixed-point per-image granule sums (one launch), float32 = row-block sums
    grans = []

    def split(st, c, nb):
        if st is None:
            return None, None
        if st.dtype == torch.int64:
            if st.dim() != 3 or st.shape[0] != nb or st.shape[2] != 2 or c % st.shape[1] or not st.is_contiguous():
                raise ValueError("per-image granule sums must be contiguous int64 [batch, C / gran, 2]")
            grans.append(c // st.shape[1])
            return st, None
        _chk(st, f32, "stats")
        return None, st
    # x1 may hold fewer images than x0 (a skip tensor computed once for both halves of a CFG pair): image b reads b % B1
    B1 = B if x1 is None else x1.shape[0]
    if B1 != B:
        if B % B1 or x1.numel() != B1 * hw * c1:
            raise ValueError("x1 must hold batch / k images of the same raster")
        if (x0_stats is not None and x0_stats.dtype != torch.int64) or (x1_stats is not None and x1_stats.dtype != torch.int64):
            x0_stats = x1_stats = None       # row-block sums cannot be shared between images: own statistics pass
    x0_sums, x0_rb = split(x0_stats, c0, B)
    x1_sums, x1_rb = split(x1_stats, c1, B1)
    one_launch = x0_sums is not None and (x1 is None or x1_sums is not None)
    if one_launch and len(set(grans)) != 1:
        raise ValueError("both sources must carry sums of the same channel granule")
    args = _lib.GroupNormArgs(x0=x0.data_ptr(), c0=c0, x1=_lib.ptr(x1), c1=c1, batch=B, hw=hw, groups=groups, eps=eps,
                              gamma=gamma.data_ptr(), beta=beta.data_ptr(), silu=int(silu),
                              out_norm=out_norm.data_ptr(), out_raw=_lib.ptr(out_raw), partials=partials.data_ptr(),
                              x0_stats=_lib.ptr(x0_rb), x1_stats=_lib.ptr(x1_rb), x0_stats_phases=x0_stats_phases,
                              x0_sums=_lib.ptr(x0_sums), x1_sums=_lib.ptr(x1_sums), sums_gran=grans[0] if grans else 0,
                              x1_batch=0 if B1 == B else B1)
    _lib.call("idb_groupnorm", C.byref(args), _lib.stream_ptr(), launches=1 if one_launch else 2)
    return out_norm, out_raw


def layernorm(x, gamma, beta, out=None, eps: float = 1e-5):
    _chk(x, f32, "x"); _chk(gamma, f32, "gamma"); _chk(beta, f32, "beta")
    c = x.shape[-1]
    rows = x.numel() // c
    if out is None:
        out = torch.empty(x.shape, dtype=bf16, device=x.device)
    _lib.call("idb_layernorm", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), rows, c, eps,
              _lib.stream_ptr())
    return out


def softmax_rows(s, scale: float, out=None):
    _chk(s, f32, "s")
    cols = s.shape[-1]
    rows = s.numel() // cols
    if out is None:
        out = torch.empty(s.shape, dtype=bf16, device=s.device)
    _lib.call("idb_softmax_rows", s.data_ptr(), out.data_ptr(), rows, cols, scale, _lib.stream_ptr())
    return out


def time_embed(timesteps, w1, b1, w2, b2, w_all, b_all, proj_out=None, scratch=None):
    """timesteps fp32 [B] -> proj_out fp32 [B, n_all] (all time_emb_proj outputs of the UNet)."""
    for t, nm in ((timesteps, "timesteps"), (w1, "w1"), (b1, "b1"), (w2, "w2"), (b2, "b2"), (w_all, "w_all"), (b_all, "b_all")):
        _chk(t, f32, nm)
    B = timesteps.shape[0]
    dim_emb, dim_sin = w1.shape
    n_all = w_all.shape[0]
    if proj_out is None:
        proj_out = torch.empty((B, n_all), dtype=f32, device=timesteps.device)
    if scratch is None:
        scratch = torch.empty((B, 2 * dim_emb + dim_sin), dtype=f32, device=timesteps.device)
    args = _lib.TimeEmbedArgs(timesteps=timesteps.data_ptr(), batch=B, dim_sin=dim_sin, dim_emb=dim_emb,
                              w1=w1.data_ptr(), b1=b1.data_ptr(), w2=w2.data_ptr(), b2=b2.data_ptr(),
                              w_all=w_all.data_ptr(), b_all=b_all.data_ptr(), n_all=n_all,
                              proj_out=proj_out.data_ptr(), scratch=scratch.data_ptr())
    _lib.call("idb_time_embed", C.byref(args), _lib.stream_ptr())
    return proj_out


def conv3x3_small_cin(x, w, bias, *, nchw: bool, out_f32=None, out_bf16=None):
    """x fp32 [B,4,H,W] (nchw) or [B,H,W,4]; w fp32 [Cout,3,3,4] -> NHWC fp32 (and/or bf16)."""
    _chk(x, f32, "x"); _chk(w, f32, "w"); _chk(bias, f32, "bias", allow_none=True)
    if nchw:
        B, cin, H, W_ = x.shape
    else:
        B, H, W_, cin = x.shape
    cout = w.shape[0]
    if out_f32 is None and out_bf16 is None:
        out_f32 = torch.empty((B, H, W_, cout), dtype=f32, device=x.device)
    _lib.call("idb_conv3x3_small_cin", x.data_ptr(), int(nchw), w.data_ptr(), _lib.ptr(bias), _lib.ptr(out_f32),
              _lib.ptr(out_bf16), B, H, W_, cin, cout, _lib.stream_ptr())
    return out_f32, out_bf16


def conv3x3_small_cout(x, w, bias, *, postprocess: bool = False, out=None):
    """x bf16 [B,H,W,Cin]; w fp32 [Cout,3,3,Cin] -> fp32 NCHW [B,Cout,H,W] (or NHWC image in [0,1])."""
    _chk(x, bf16, "x"); _chk(w, f32, "w"); _chk(bias, f32, "bias", allow_none=True)
    B, H, W_, cin = x.shape
    cout = w.shape[0]
    if out is None:
        out = torch.empty((B, H, W_, cout) if postprocess else (B, cout, H, W_), dtype=f32, device=x.device)
    _lib.call("idb_conv3x3_small_cout", x.data_ptr(), w.data_ptr(), _lib.ptr(bias), out.data_ptr(), int(postprocess),
              B, H, W_, cin, cout, _lib.stream_ptr())
    return out


def upsample2x(x, out=None):
    _chk(x, f32, "x")
    B, H, W_, c = x.shape
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W_, c), dtype=bf16, device=x.device)
    _lib.call("idb_upsample2x", x.data_ptr(), out.data_ptr(), B, H, W_, c, _lib.stream_ptr())
    return out


def cast_bf16(x, out=None):
    _chk(x, f32, "x")
    if out is None:
        out = torch.empty(x.shape, dtype=bf16, device=x.device)
    _lib.call("idb_cast_bf16", x.data_ptr(), out.data_ptr(), x.numel(), _lib.stream_ptr())
    return out


def vae_latent_prep(z, w, bias, inv_scaling: float, out=None):
    _chk(z, f32, "z"); _chk(w, f32, "w"); _chk(bias, f32, "bias")
    B, c, H, W_ = z.shape
    if out is None:
        out = torch.empty((B, H, W_, 4), dtype=f32, device=z.device)
    _lib.call("idb_vae_latent_prep", z.data_ptr(), w.data_ptr(), bias.data_ptr(), inv_scaling, out.data_ptr(), B, H * W_,
              _lib.stream_ptr())
    return out


def latent_operand(x, out=None):
    """x fp32 NCHW [B,4,H,W] -> bf16 NHWC [B,H,W,64] (hi | lo | hi split of the 4 channels) for the tensor-core conv_in."""
    _chk(x, f32, "x")
    B, c, H, W_ = x.shape
    if c != 4:
        raise ValueError("latent_operand expects 4 channels")
    if out is None:
        out = torch.empty((B, H, W_, 64), dtype=bf16, device=x.device)
    _chk(out, bf16, "out")
    _lib.call("idb_latent_operand", x.data_ptr(), out.data_ptr(), B, H * W_, _lib.stream_ptr())
    return out


def cfg_ddpm_step(eps2, x, noise, coef, *, guidance_scale: float, use_cfg: bool, v_prediction: bool = False,
                  x_prev=None, x0_out=None):
    """eps2 fp32 [2n or n, ...], x fp32 [n, ...], coef fp32[5] on device."""
    _chk(eps2, f32, "eps2"); _chk(x, f32, "x"); _chk(noise, f32, "noise", allow_none=True); _chk(coef, f32, "coef")
    if x_prev is None:
        x_prev = torch.empty_like(x)
    _lib.call("idb_cfg_ddpm_step", eps2.data_ptr(), x.data_ptr(), _lib.ptr(noise), coef.data_ptr(), guidance_scale,
              int(use_cfg), int(v_prediction), x_prev.data_ptr(), _lib.ptr(x0_out), x.numel(), _lib.stream_ptr())
    return x_prev


def channel_affine(x, scale=None, shift=None, *, stride: int = 1, out=None, half: bool = False):
    """x fp32 NHWC [B,H,W,C] -> bf16 [B,H/s,W/s,C] = x[:, ::s, ::s] * scale + shift (eval BatchNorm2d / stride-s sampling)."""
    _chk(x, f32, "x"); _chk(scale, f32, "scale", allow_none=True); _chk(shift, f32, "shift", allow_none=True)
    B, H, W_, Cc = x.shape
    if out is None:
        out = torch.empty((B, H // stride, W_ // stride, Cc), dtype=f16 if half else bf16, device=x.device)
    _chk(out, f16 if half else bf16, "out")
    _lib.call("idb_channel_affine", x.data_ptr(), _lib.ptr(scale), _lib.ptr(shift), out.data_ptr(), int(half), B, H, W_, Cc,
              stride, _lib.stream_ptr())
    return out


def crop_resize_norm(images, bbox, *, size: int = 112, c_pad: int = 64, out=None, half: bool = False):
    """images fp32 NHWC [n,H,W,3] in [0,1], bbox int32 [n,4] (x0,y0,x1,y1) -> bf16 [n,size,size,c_pad] ArcFace input."""
    _chk(images, f32, "images"); _chk(bbox, torch.int32, "bbox")
    n, H, W_, c = images.shape
    if c != 3 or tuple(bbox.shape) != (n, 4):
        raise ValueError("images must be [n,H,W,3] and bbox [n,4]")
    if out is None:
        out = torch.empty((n, size, size, c_pad), dtype=f16 if half else bf16, device=images.device)
    _chk(out, f16 if half else bf16, "out")
    _lib.call("idb_crop_resize_norm", images.data_ptr(), bbox.data_ptr(), out.data_ptr(), int(half), n, H, W_, size, c_pad,
              _lib.stream_ptr())
    return out
