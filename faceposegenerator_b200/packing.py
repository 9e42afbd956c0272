"""One-time weight repacking into the layouts the kernels consume (done at load, into
caller-owned tensors; base weights are never modified afterwards).

 - conv weights [Cout, Cin, kh, kw] (diffusers) -> K-major [Cout, kh*kw*Cin], tap-major /
   channel-minor, optionally with the 1x1 conv_shortcut appended on K;
 - GEGLU `ff.net.0.proj` rows interleaved in 16-blocks [a | g] so `a * gelu(g)` is local
   to one 32-column epilogue chunk;
 - LoRA A zero-padded to 16 rows per adapter (16 extra UMMA N-columns), B*scale as bf16
   [N, 64] K-major operand rows (the up-projection is one more UMMA per tile).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

bf16, f32 = torch.bfloat16, torch.float32


def pack_conv_weight(w: torch.Tensor, shortcut: Optional[torch.Tensor] = None, device=None) -> torch.Tensor:
    """[Cout, Cin, kh, kw] (+ optional [Cout, Cs, 1, 1]) -> bf16 [Cout, kh*kw*Cin (+Cs)]."""
    p = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    if shortcut is not None:
        p = torch.cat([p, shortcut.reshape(shortcut.shape[0], -1)], dim=1)
    return p.to(device=device, dtype=bf16).contiguous()


def pack_upsample_phase_weights(w: torch.Tensor, device=None, stacked: bool = False):
    """Upsample2D = nearest-2x + conv3x3(pad 1).  Output pixel (2i+a, 2j+c) only sees the 2x2 low-resolution
    neighbourhood rows {i+a-1, i+a}, cols {j+c-1, j+c}; the 3x3 taps that land on the same input pixel are summed in
    fp32: a = 0: rows (w0 | w1+w2), a = 1: rows (w0+w1 | w2); same for columns.
    Returns w_ph[a][c]: bf16 [Cout, 4*Cin] (tap = 2u + v major, channel minor) for idb_gemm_conv(IDB_A_2X2) with
    tap offsets (a - 1, c - 1); `stacked=True`: (w_all, w_ph) where w_all is ONE bf16 [4*Cout, 4*Cin] tensor, phase 2a + c
    major (the operand of IDB_EPI_PHASES4), and w_ph[a][c] are row-slice views of it."""
    w = w.float()                                        # [Cout, Cin, 3, 3]
    rows = ((w[:, :, 0:1], w[:, :, 1:2] + w[:, :, 2:3]), (w[:, :, 0:1] + w[:, :, 1:2], w[:, :, 2:3]))
    out = []
    for a in range(2):
        r = torch.cat(rows[a], dim=2)                    # [Cout, Cin, 2, 3]
        cols = ((r[..., 0:1], r[..., 1:2] + r[..., 2:3]), (r[..., 0:1] + r[..., 1:2], r[..., 2:3]))
        out.append([torch.cat(cols[c], dim=3).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
                    .to(device=device, dtype=bf16).contiguous() for c in range(2)])
    if not stacked:
        return out
    cout = w.shape[0]
    w_all = torch.cat([out[a][c] for a in range(2) for c in range(2)], dim=0).contiguous()
    return w_all, [[w_all[(2 * a + c) * cout:(2 * a + c + 1) * cout] for c in range(2)] for a in range(2)]


def pack_edge_conv_weight(w: torch.Tensor, device=None) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> fp32 [Cout, 3, 3, Cin] for the tiny-Cin / tiny-Cout SIMT kernels."""
    return w.permute(0, 2, 3, 1).to(device=device, dtype=f32).contiguous()


def interleave_geglu(w: torch.Tensor, bias: Optional[torch.Tensor]):
    """w [2*H, K] = [value rows | gate rows] -> rows ordered [v(16) g(16) v(16) g(16) ...]."""
    H = w.shape[0] // 2
    assert H % 16 == 0
    idx = torch.arange(2 * H, device=w.device).view(2, H // 16, 16).permute(1, 0, 2).reshape(-1)
    wi = w.index_select(0, idx).contiguous()
    bi = bias.index_select(0, idx).contiguous() if bias is not None else None
    return wi, bi


LORA_UP_COLS = 64   # bf16 columns of the packed up-projection operand (128-byte rows)


def pack_lora(adapters: Sequence[Optional[Tuple[torch.Tensor, torch.Tensor, float]]], device=None,
              seg_n: Optional[int] = None, k: Optional[int] = None):
    """adapters: one (down [r, K], up [seg_n, r], scale) per N-segment of a fused projection
    (None = segment without adapter).  Returns (down bf16 [nseg*16, K], up bf16 [nseg*seg_n, 64]):
    `up` row n holds B[n, :] * scale in its first `rank` columns (one 128-byte K-major operand row)."""
    live = [a for a in adapters if a is not None]
    if not live:
        return None, None
    r = max(a[0].shape[0] for a in live)
    if r > 16:
        raise ValueError("fused LoRA supports rank <= 16")
    rank_pad = (r + 3) // 4 * 4
    k = k or live[0][0].shape[1]
    seg_n = seg_n or live[0][1].shape[0]
    src = live[0][0].device      # adapters that already live on the GPU (training) are packed there, without a host round trip
    down = torch.zeros((len(adapters) * 16, k), dtype=f32, device=src)
    up = torch.zeros((len(adapters) * seg_n, LORA_UP_COLS), dtype=f32, device=src)
    for s, a in enumerate(adapters):
        if a is None:
            continue
        d, u, scale = a
        down[s * 16:s * 16 + d.shape[0]] = d.detach().float().to(src)
        up[s * seg_n:(s + 1) * seg_n, :u.shape[1]] = u.detach().float().to(src) * float(scale)
    return down.to(device=device, dtype=bf16).contiguous(), up.to(device=device, dtype=bf16).contiguous()
