"""T1 op parity: every C-ABI kernel vs the fp32 torch op it replaces (same bf16-rounded
operands, fp32 math, TF32 off).  Tolerances: fp32 outputs rel-L2 <= 2e-5 (accumulation
order only), bf16 outputs <= 4e-3 (one bf16 rounding)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rb(t):  # bf16 round trip
    return t.to(torch.bfloat16)


def fixed_sums(t):  # int64 [.., 2] fixed point (sum * 2^32, sum of squares * 2^24) -> float64 pair
    return t[..., 0].double() / 2.0 ** 32, t[..., 1].double() / 2.0 ** 24


def pack_conv_w(w):  # [Cout, Cin, 3, 3] -> [Cout, 9*Cin] tap-major
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


@pytest.fixture(scope="module")
def ops(cuda_dev):
    from faceposegenerator_b200 import ops
    return ops


# ------------------------------------------------------------------ Linear family
@pytest.mark.parametrize("M,K,N", [(256, 128, 128), (4096, 320, 320), (1024, 640, 2560), (154, 1024, 640),
                                   (8192, 1280, 1280), (300, 64, 32), (2048, 5120, 1280)])
def test_linear_plain(ops, cuda_dev, M, K, N):
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = rb(torch.randn(M, K, device=cuda_dev, generator=g))
    w = rb(torch.randn(N, K, device=cuda_dev, generator=g) / math.sqrt(K))
    o32, o16 = ops.gemm_conv(x, w, want_f32=True, want_bf16=True)
    ref = x.float() @ w.float().t()
    torch.cuda.synchronize()
    assert rel(o32, ref) < 2e-5, rel(o32, ref)
    assert rel(o16.float(), ref) < 4e-3


def test_linear_bias_residual_rowvec(ops, cuda_dev):
    M, K, N = 2048, 640, 640
    g = torch.Generator(device="cuda").manual_seed(1)
    x = rb(torch.randn(M, K, device=cuda_dev, generator=g))
    w = rb(torch.randn(N, K, device=cuda_dev, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=cuda_dev, generator=g)
    res = torch.randn(M, N, device=cuda_dev, generator=g)
    o32, o16 = ops.gemm_conv(x, w, bias=bias, residual=res, want_f32=True, want_bf16=True)
    ref = x.float() @ w.float().t() + bias + res
    assert rel(o32, ref) < 2e-5
    assert rel(o16.float(), ref) < 4e-3


@pytest.mark.parametrize("M,K,N", [(4096, 320, 320), (154, 1024, 640), (8192, 1280, 1280), (32768, 320, 320),
                                   (2048, 640, 2560), (300, 64, 32), (100, 320, 320)])
@pytest.mark.parametrize("out", ["f32", "bf16"])
def test_linear_residual_single_output(ops, cuda_dev, M, K, N, out):
    """One output + fp32 residual: the epilogue prefetches the residual by TMA into its staging buffers
    (the path every residual-stream GEMM of the UNet takes)."""
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = rb(torch.randn(M, K, device=cuda_dev, generator=g))
    w = rb(torch.randn(N, K, device=cuda_dev, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=cuda_dev, generator=g)
    res = torch.randn(M, N, device=cuda_dev, generator=g)
    ref = x.float() @ w.float().t() + bias + res
    for _ in range(2):   # twice: staging-buffer / barrier phases must be clean across launches
        o32, o16 = ops.gemm_conv(x, w, bias=bias, residual=res, want_f32=out == "f32", want_bf16=out == "bf16")
        if out == "f32":
            assert o16 is None and rel(o32, ref) < 2e-5, rel(o32, ref)
        else:
            assert o32 is None and rel(o16.float(), ref) < 4e-3


def test_linear_geglu(ops, cuda_dev):
    M, K, C4 = 1024, 320, 1280
    g = torch.Generator(device="cuda").manual_seed(2)
    x = rb(torch.randn(M, K, device=cuda_dev, generator=g))
    w = rb(torch.randn(2 * C4, K, device=cuda_dev, generator=g) / math.sqrt(K))
    bias = 0.1 * torch.randn(2 * C4, device=cuda_dev, generator=g)
    from faceposegenerator_b200.packing import interleave_geglu
    wi, bi = interleave_geglu(w, bias)
    _, o16 = ops.gemm_conv(x, wi, bias=bi, geglu=True, want_bf16=True)
    y = x.float() @ w.float().t() + bias
    a, gate = y.chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    assert o16.shape == (M, C4)
    assert rel(o16.float(), ref) < 4e-3


@pytest.mark.parametrize("C,Kin,nseg", [(320, 320, 3), (640, 1024, 2), (1280, 1280, 1)])
def test_linear_fused_lora(ops, cuda_dev, C, Kin, nseg):
    """Fused q/k/v (or k/v, or single) projection with one rank-4 adapter per segment:
    y_s = x W_s^T + (x A_s^T) B_s^T -- peft lora.Linear, unmerged."""
    from faceposegenerator_b200.packing import pack_lora
    M, r = 1024, 4
    g = torch.Generator(device="cuda").manual_seed(C)
    x = rb(torch.randn(M, Kin, device=cuda_dev, generator=g))
    ws = [rb(torch.randn(C, Kin, device=cuda_dev, generator=g) / math.sqrt(Kin)) for _ in range(nseg)]
    downs = [torch.randn(r, Kin, device=cuda_dev, generator=g) / r for _ in range(nseg)]
    ups = [torch.randn(C, r, device=cuda_dev, generator=g) * 0.05 for _ in range(nseg)]
    bias = torch.randn(nseg * C, device=cuda_dev, generator=g)
    w = torch.cat(ws, 0).contiguous()
    ld, lu = pack_lora(list(zip(downs, ups, [1.0] * nseg)), device=cuda_dev)
    o32, _ = ops.gemm_conv(x, w, bias=bias, lora_down=ld, lora_up=lu, lora_seg_n=C, want_f32=True)
    refs, exact = [], []
    for s in range(nseg):
        base = x.float() @ ws[s].float().t()
        t = x.float() @ rb(downs[s]).float().t()
        exact.append(base + t @ ups[s].t())
        # the kernel's arithmetic: x A^T accumulated in fp32, rounded to bf16, times bf16(B * scale), fp32 accumulate
        refs.append(base + rb(t).float() @ rb(ups[s]).float().t())
    ref = torch.cat(refs, 1) + bias
    assert rel(o32, ref) < 1e-4, rel(o32, ref)   # (a few T elements sit on bf16 rounding boundaries)
    assert rel(o32, torch.cat(exact, 1) + bias) < 3e-3   # bf16 rounding of the rank-r factors only (this adapter is as large as the base)
    # the adapter really contributes
    assert rel(o32, torch.cat([x.float() @ wsi.float().t() for wsi in ws], 1) + bias) > 1e-3
    # bf16 output + residual (attention out-projection path), and a single-M-tile problem (1-CTA kernel)
    res = torch.randn(M, nseg * C, device=cuda_dev, generator=g)
    o32r, _ = ops.gemm_conv(x, w, bias=bias, residual=res, lora_down=ld, lora_up=lu, lora_seg_n=C, want_f32=True)
    assert rel(o32r, ref + res) < 1e-4
    _, o16 = ops.gemm_conv(x, w, bias=bias, lora_down=ld, lora_up=lu, lora_seg_n=C, want_bf16=True)
    assert rel(o16.float(), ref) < 4e-3
    o32s, _ = ops.gemm_conv(x[:100].contiguous(), w, bias=bias, lora_down=ld, lora_up=lu, lora_seg_n=C, want_f32=True)
    assert rel(o32s, ref[:100]) < 1e-4


@pytest.mark.parametrize("M,K,N,splits", [(512, 11520, 1280, 4), (128, 23040, 1280, 8), (2048, 640, 640, 3)])
def test_linear_split_k(ops, cuda_dev, M, K, N, splits):
    g = torch.Generator(device="cuda").manual_seed(7)
    x = rb(torch.randn(M, K, device=cuda_dev, generator=g))
    w = rb(torch.randn(N, K, device=cuda_dev, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=cuda_dev, generator=g)
    res = torch.randn(M, N, device=cuda_dev, generator=g)
    o32, o16 = ops.gemm_conv(x, w, bias=bias, residual=res, want_f32=True, want_bf16=True, k_splits=splits)
    ref = x.float() @ w.float().t() + bias + res
    assert rel(o32, ref) < 2e-5
    assert rel(o16.float(), ref) < 4e-3


# ------------------------------------------------------------------ convolutions
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 128), (2, 64, 64, 320, 320), (3, 8, 8, 1280, 1280), (8, 64, 64, 320, 320),
                                            (8, 32, 32, 640, 640),
                                            (1, 32, 32, 640, 640), (2, 24, 24, 64, 128), (1, 128, 128, 128, 128),
                                            (1, 12, 20, 64, 64)])
def test_conv3x3(ops, cuda_dev, B, H, W, Cin, Cout):
    g = torch.Generator(device="cuda").manual_seed(B * H + Cin)
    x = rb(torch.randn(B, H, W, Cin, device=cuda_dev, generator=g))
    w = rb(torch.randn(Cout, Cin, 3, 3, device=cuda_dev, generator=g) / math.sqrt(9 * Cin))
    bias = torch.randn(Cout, device=cuda_dev, generator=g)
    rowvec = torch.randn(B, Cout, device=cuda_dev, generator=g)
    o32, _ = ops.gemm_conv(x, pack_conv_w(w), mode=ops.A_3X3, bias=bias, rowvec=rowvec, want_f32=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1) + rowvec[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    assert rel(o32, ref) < 2e-5, rel(o32, ref)


@pytest.mark.parametrize("B,H,W,Cin,Cout,C1,sk", [(8, 16, 16, 1280, 1280, 0, 1), (8, 16, 16, 1280, 1280, 640, 1), (4, 16, 16, 1280, 1280, 0, 1),
                                                  (8, 16, 16, 128, 640, 0, 0),    # K too short to share: plain tiles
                                                  (6, 16, 16, 2560, 1280, 0, 1), (2, 16, 16, 1280, 1280, 0, 0)])
def test_conv3x3_stream_k(ops, cuda_dev, B, H, W, Cin, Cout, C1, sk):
    """One-wave conv layers (16x16 latents) take the per-image stream-K schedule: dual-N tiles whose K range is shared by
    several CTA pairs, partial tiles summed by the owning pair in a fixed order.  Against fp32 conv2d (+ 1x1 shortcut
    segment, bias, time-embedding row vector, residual, per-image sums); bit-reproducible across calls (the arrival
    counters re-arm themselves) and independent of where an image sits in the batch."""
    from faceposegenerator_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(B * H + Cin + C1)
    x = rb(torch.randn(B, H, W, Cin, device=cuda_dev, generator=g))
    w = rb(torch.randn(Cout, Cin, 3, 3, device=cuda_dev, generator=g) / math.sqrt(9 * Cin))
    x1 = rb(torch.randn(B, H, W, C1, device=cuda_dev, generator=g)) if C1 else None
    w1 = rb(torch.randn(Cout, C1, device=cuda_dev, generator=g) / math.sqrt(C1)) if C1 else None
    wk = torch.cat([pack_conv_w(w), w1], 1).contiguous() if C1 else pack_conv_w(w)
    bias = torch.randn(Cout, device=cuda_dev, generator=g)
    rowvec = torch.randn(B, Cout, device=cuda_dev, generator=g)
    res = torch.randn(B * H * W, Cout, device=cuda_dev, generator=g)
    ws = torch.empty(24 << 20, dtype=torch.float32, device=cuda_dev)

    def run(xa, x1a, rv, rs):
        return ops.gemm_conv(xa, wk, mode=ops.A_3X3, a1=x1a, bias=bias, rowvec=rv, residual=rs, want_f32=True, want_stats=True,
                             k_splits=0, workspace=ws, stats_gran=10)
    n0 = _lib.load().idb_stream_k_launch_count()
    o, _, sm = run(x, x1, rowvec, res)
    took_sk = _lib.load().idb_stream_k_launch_count() - n0
    if os.environ.get("IDB_GEMM_SK", "1") == "1":
        assert took_sk == sk, "schedule selection changed: update the test's expectation"
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
    if C1:
        ref = ref + F.conv2d(x1.float().permute(0, 3, 1, 2), w1.float()[:, :, None, None])
    ref = (ref + rowvec[:, :, None, None]).permute(0, 2, 3, 1).reshape(B * H * W, Cout) + res
    assert rel(o, ref) < 2e-5, rel(o, ref)
    flat = o.view(B, H * W, Cout // 10, 10).double().permute(0, 1, 3, 2).reshape(B, H * W * 10, Cout // 10)
    assert rel(fixed_sums(sm)[0], flat.sum(1)) < 1e-6
    for _ in range(3):
        o2, _, sm2 = run(x, x1, rowvec, res)
        assert torch.equal(o2, o) and torch.equal(sm2, sm)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(cuda_dev)
    op, _, smp = run(x[perm].contiguous(), x1[perm].contiguous() if C1 else None, rowvec[perm].contiguous(),
                     res.view(B, H * W, Cout)[perm].reshape(B * H * W, Cout).contiguous())
    assert torch.equal(op.view(B, H * W, Cout), o.view(B, H * W, Cout)[perm]) and torch.equal(smp, sm[perm])


@pytest.mark.parametrize("B,H,W,C", [(2, 64, 64, 320), (2, 16, 16, 1280), (1, 8, 8, 64)])
def test_conv3x3_stride2(ops, cuda_dev, B, H, W, C):
    g = torch.Generator(device="cuda").manual_seed(H)
    x = rb(torch.randn(B, H, W, C, device=cuda_dev, generator=g))
    w = rb(torch.randn(C, C, 3, 3, device=cuda_dev, generator=g) / math.sqrt(9 * C))
    bias = torch.randn(C, device=cuda_dev, generator=g)
    o32, _ = ops.gemm_conv(x, pack_conv_w(w), mode=ops.A_3X3_S2, bias=bias, want_f32=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=2, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, C)
    assert rel(o32, ref) < 2e-5, rel(o32, ref)


@pytest.mark.parametrize("B,H,W,C", [(2, 64, 64, 128), (1, 32, 32, 256), (3, 16, 16, 512)])
def test_conv3x3_stride2_asymmetric(ops, cuda_dev, B, H, W, C):
    """VAE encoder Downsample2D: F.pad(x, (0, 1, 0, 1)) then conv3x3 stride 2 padding 0."""
    g = torch.Generator(device="cuda").manual_seed(H + 1)
    x = rb(torch.randn(B, H, W, C, device=cuda_dev, generator=g))
    w = rb(torch.randn(C, C, 3, 3, device=cuda_dev, generator=g) / math.sqrt(9 * C))
    bias = torch.randn(C, device=cuda_dev, generator=g)
    o32, _ = ops.gemm_conv(x, pack_conv_w(w), mode=ops.A_3X3_S2_ASYM, bias=bias, want_f32=True)
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), w.float(), bias, stride=2, padding=0)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, C)
    assert rel(o32, ref) < 2e-5, rel(o32, ref)


def test_linear_gelu_epilogue(ops, cuda_dev):
    """CLIP text MLP fc1: gelu_erf(x W^T + b) in the GEMM epilogue."""
    M, K, N = 616, 1024, 4096
    g = torch.Generator(device="cuda").manual_seed(12)
    x = rb(torch.randn(M, K, device=cuda_dev, generator=g))
    w = rb(torch.randn(N, K, device=cuda_dev, generator=g) / math.sqrt(K))
    bias = torch.randn(N, device=cuda_dev, generator=g)
    _, o16 = ops.gemm_conv(x, w, bias=bias, gelu=True, want_bf16=True)
    ref = F.gelu(x.float() @ w.float().t() + bias)
    assert rel(o16.float(), ref) < 4e-3


def test_attention_causal_short(ops, cuda_dev):
    """CLIP text tower self-attention: 77 tokens, causal mask, fused q/k/v buffer."""
    B, heads, T = 3, 16, 77
    C = heads * 64
    g = torch.Generator(device="cuda").manual_seed(21)
    qkv = rb(torch.randn(B * T, 3 * C, device=cuda_dev, generator=g))
    out = ops.attention(qkv, qkv, qkv, batch=B, heads=heads, t_q=T, t_kv=T, scale=0.125, col0_q=0, col0_k=C, col0_v=2 * C,
                        causal=True)
    q, k, v = qkv.float().view(B, T, 3, heads, 64).unbind(2)
    ref = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
    ref = ref.transpose(1, 2).reshape(B * T, C)
    assert rel(out.float(), ref) < 6e-3


@pytest.mark.parametrize("B,H,W,C", [(8, 32, 32, 640), (8, 16, 16, 1280), (2, 16, 16, 128),
                                     (4, 8, 8, 1280)])   # 16x16 output: the GroupNorm finalizes the phased statistics inside its apply kernel
def test_upsample_conv_four_phase(ops, cuda_dev, B, H, W, C):
    """Upsample2D (nearest 2x + conv3x3) as four 2x2 convolutions on the low-resolution tensor, written phase by phase
    into the full-resolution output; the GroupNorm that follows consumes the phased epilogue statistics."""
    from faceposegenerator_b200.packing import pack_upsample_phase_weights
    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = torch.randn(B, H, W, C, device=cuda_dev, generator=g)
    w = torch.randn(C, C, 3, 3, device=cuda_dev, generator=g) / math.sqrt(9 * C)
    bias = torch.randn(C, device=cuda_dev, generator=g)
    wph = pack_upsample_phase_weights(w, device=cuda_dev)
    xb = ops.cast_bf16(x)
    out = torch.full((B, 2 * H, 2 * W, C), float("nan"), device=cuda_dev)
    st = torch.empty((4, B * H * W // 32, C, 2), device=cuda_dev)
    for a in range(2):
        for c in range(2):
            ops.gemm_conv(xb, wph[a][c], mode=ops.A_2X2, bias=bias, out_f32=out, stats=st, tap_off=(a - 1, c - 1), out_phase=(a, c))
    up = F.interpolate(xb.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w, bias, padding=1).permute(0, 2, 3, 1)
    assert not torch.isnan(out).any()
    assert rel(out, ref) < 3e-3, rel(out, ref)          # the pre-summed taps are rounded to bf16 once more than the 3x3 weights
    gamma = torch.randn(C, device=cuda_dev, generator=g)
    beta = torch.randn(C, device=cuda_dev, generator=g)
    y_ph, _ = ops.groupnorm(out, gamma, beta, groups=32, eps=1e-5, silu=True, x0_stats=st, x0_stats_phases=4)
    y_pl, _ = ops.groupnorm(out, gamma, beta, groups=32, eps=1e-5, silu=True)
    assert rel(y_ph.float(), y_pl.float()) < 2e-3
    # the same four calls also accumulating the fixed-point per-image channel sums (zeroed once, before the first phase)
    assert ops.image_sums_supported(B, H * W, C, phased=True)
    out2 = torch.empty_like(out)
    sums = torch.zeros((B, C, 2), dtype=torch.int64, device=cuda_dev)
    for a in range(2):
        for c in range(2):
            ops.gemm_conv(xb, wph[a][c], mode=ops.A_2X2, bias=bias, out_f32=out2, sums=sums, tap_off=(a - 1, c - 1),
                          out_phase=(a, c))
    assert torch.equal(out2, out)
    flat = out.view(B, 4 * H * W, C).double()
    assert rel(fixed_sums(sums)[0], flat.sum(1)) < 1e-6 and rel(fixed_sums(sums)[1], (flat ** 2).sum(1)) < 1e-6
    y_sm, _ = ops.groupnorm(out, gamma, beta, groups=32, eps=1e-5, silu=True, x0_stats=sums)
    assert rel(y_sm.float(), y_pl.float()) < 2e-3
    # all four phases in ONE call (IDB_EPI_PHASES4: phase weights stacked on N): the same bits, statistics and sums
    w_all, views = pack_upsample_phase_weights(w, device=cuda_dev, stacked=True)
    assert tuple(w_all.shape) == (4 * C, 4 * C) and all(torch.equal(views[a][c], wph[a][c]) for a in range(2) for c in range(2))
    out3 = torch.full_like(out, float("nan"))
    st3 = torch.full_like(st, float("nan"))
    sums3 = torch.zeros_like(sums)
    ops.gemm_conv(xb, w_all, mode=ops.A_2X2, bias=bias, out_f32=out3, stats=st3, sums=sums3, phases4=True)
    assert torch.equal(out3, out) and torch.equal(st3, st) and torch.equal(sums3, sums)


@pytest.mark.parametrize("B,H,W,Cin,N,gran", [(8, 64, 64, 64, 320, 10), (8, 64, 64, 64, 320, 1),
                                              (3, 8, 8, 64, 1280, 40),      # odd batch with two images per tile: padding row blocks
                                              (1, 8, 8, 128, 640, 20), (5, 16, 16, 64, 1280, 10), (2, 32, 32, 64, 640, 4),
                                              (8, 8, 8, 1280, 1280, 10),    # K = 11520, M = 512: split-K + fused finalize / sums kernel
                                              (8, 16, 16, 1280, 1280, 1), (4, 64, 64, 64, 128, 4)])
def test_gemm_image_sums(ops, cuda_dev, B, H, W, Cin, N, gran):
    """`stats_image_sums`: per-image channel (sum, sum of squares) of the fp32 output in 64-bit fixed point, accumulated by
    integer atomics in the conv GEMM's epilogue (exact and order-independent) or written by the fused split-K finalize --
    against float64 sums of the output tensor, and bit-reproducible across calls."""
    g = torch.Generator(device="cuda").manual_seed(B * H + N)
    x = rb(torch.randn(B, H, W, Cin, device=cuda_dev, generator=g))
    w = rb(torch.randn(N, Cin, 3, 3, device=cuda_dev, generator=g) / math.sqrt(9 * Cin))
    bias = torch.randn(N, device=cuda_dev, generator=g)
    ws = torch.empty(24 << 20, dtype=torch.float32, device=cuda_dev)
    runs = []
    for _ in range(3):
        o, _, sm = ops.gemm_conv(x, pack_conv_w(w), mode=ops.A_3X3, bias=bias, want_f32=True, want_stats=True, k_splits=0, workspace=ws,
                                 stats_gran=gran)
        assert sm.dtype == torch.int64 and tuple(sm.shape) == (B, N // gran, 2)
        runs.append((o.clone(), sm.clone()))
    flat = runs[0][0].view(B, H * W, N // gran, gran).double().permute(0, 1, 3, 2).reshape(B, H * W * gran, N // gran)
    assert rel(fixed_sums(runs[0][1])[0], flat.sum(1)) < 1e-6
    assert rel(fixed_sums(runs[0][1])[1], (flat ** 2).sum(1)) < 1e-6
    for o, sm in runs[1:]:
        assert torch.equal(o, runs[0][0]) and torch.equal(sm, runs[0][1])


def test_linear_image_sums_over_tokens(ops, cuda_dev):
    """A Linear over a token matrix [B * T, K] (Transformer2DModel.proj_out + residual) hands the GroupNorm of the next
    ResnetBlock2D its per-image sums: `stats_hw` = tokens per image."""
    B, T, C = 4, 1024, 640
    g = torch.Generator(device="cuda").manual_seed(3)
    a = rb(torch.randn(B * T, C, device=cuda_dev, generator=g))
    w = rb(torch.randn(C, C, device=cuda_dev, generator=g) / math.sqrt(C))
    res = torch.randn(B * T, C, device=cuda_dev, generator=g)
    o, _, sm = ops.gemm_conv(a, w, residual=res, want_f32=True, want_stats=True, stats_hw=T)
    assert sm.dtype == torch.int64 and tuple(sm.shape) == (B, C, 2)
    flat = o.view(B, T, C).double()
    assert rel(fixed_sums(sm)[0], flat.sum(1)) < 1e-6 and rel(fixed_sums(sm)[1], (flat ** 2).sum(1)) < 1e-6
    gamma = 1 + 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    beta = 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    y, _ = ops.groupnorm(o.view(B, T, C), gamma, beta, groups=32, eps=1e-5, silu=True, x0_stats=sm)
    ref = F.silu(F.group_norm(o.view(B, T, C).permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1)
    assert rel(y.float(), ref) < 4e-3


def test_conv3x3_plus_shortcut_segment(ops, cuda_dev):
    """ResnetBlock2D tail: conv2(h) + conv_shortcut(x) + x-independent bias, as ONE GEMM
    whose K axis is [9*Cout | Cin]."""
    B, H, W, Cin, Cout = 2, 32, 32, 960, 640
    g = torch.Generator(device="cuda").manual_seed(5)
    h = rb(torch.randn(B, H, W, Cout, device=cuda_dev, generator=g))
    x = rb(torch.randn(B, H, W, Cin, device=cuda_dev, generator=g))
    w2 = rb(torch.randn(Cout, Cout, 3, 3, device=cuda_dev, generator=g) / math.sqrt(9 * Cout))
    ws = rb(torch.randn(Cout, Cin, 1, 1, device=cuda_dev, generator=g) / math.sqrt(Cin))
    b2 = torch.randn(Cout, device=cuda_dev, generator=g)
    wcat = torch.cat([pack_conv_w(w2), ws.reshape(Cout, Cin)], 1).contiguous()
    o32, _ = ops.gemm_conv(h, wcat, mode=ops.A_3X3, a1=x, bias=b2, want_f32=True)
    ref = F.conv2d(h.float().permute(0, 3, 1, 2), w2.float(), b2, padding=1) + \
        F.conv2d(x.float().permute(0, 3, 1, 2), ws.float())
    ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    assert rel(o32, ref) < 2e-5, rel(o32, ref)


# ------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,heads,Tq,Tkv", [(2, 5, 4096, 4096), (2, 10, 1024, 1024), (2, 20, 256, 256), (3, 20, 64, 64),
                                            (2, 5, 4096, 77), (1, 20, 64, 77), (1, 2, 200, 333), (8, 10, 1024, 77),
                                            (8, 20, 256, 77), (2, 3, 300, 50), (2, 3, 700, 96), (1, 2, 130, 16),
                                            # more work items than SMs: the persistent row-split kernel walks several items per CTA
                                            (8, 10, 1024, 1024), (4, 5, 2304, 1000), (6, 5, 1100, 1024)])
def test_attention(ops, cuda_dev, B, heads, Tq, Tkv):
    C = heads * 64
    g = torch.Generator(device="cuda").manual_seed(Tq + Tkv)
    # q/k/v live in fused buffers, as the pipeline lays them out
    qkv = rb(torch.randn(B * Tq, 3 * C, device=cuda_dev, generator=g))
    kv = rb(torch.randn(B * Tkv, 2 * C, device=cuda_dev, generator=g))
    if Tq == Tkv:
        out = ops.attention(qkv, qkv, qkv, batch=B, heads=heads, t_q=Tq, t_kv=Tkv, scale=0.125,
                            col0_q=0, col0_k=C, col0_v=2 * C)
        q, k, v = qkv.float().view(B, Tq, 3, heads, 64).unbind(2)
    else:
        out = ops.attention(qkv, kv, kv, batch=B, heads=heads, t_q=Tq, t_kv=Tkv, scale=0.125,
                            col0_q=0, col0_k=0, col0_v=C)
        q = qkv.float().view(B, Tq, 3, heads, 64)[:, :, 0]
        k, v = kv.float().view(B, Tkv, 2, heads, 64).unbind(2)
    ref = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    ref = ref.transpose(1, 2).reshape(B * Tq, C)
    e = rel(out.float(), ref)
    assert e < 6e-3, e


# ------------------------------------------------------------------ norms
@pytest.mark.parametrize("B,HW,C0,C1,silu,eps", [(2, 4096, 320, 0, True, 1e-5), (2, 1024, 640, 320, True, 1e-5),
                                                (3, 64, 1280, 1280, True, 1e-5), (2, 256, 1280, 640, False, 1e-6),
                                                (1, 16384, 128, 0, True, 1e-6), (2, 4096, 320, 0, False, 1e-6)])
def test_groupnorm(ops, cuda_dev, B, HW, C0, C1, silu, eps):
    g = torch.Generator(device="cuda").manual_seed(HW + C0)
    x0 = torch.randn(B, HW, C0, device=cuda_dev, generator=g) * 2 + 0.5
    x1 = torch.randn(B, HW, C1, device=cuda_dev, generator=g) if C1 else None
    C = C0 + C1
    gamma = 1 + 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    beta = 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    yn, yr = ops.groupnorm(x0, gamma, beta, groups=32, eps=eps, silu=silu, x1=x1, want_raw=True)
    xc = torch.cat([x0, x1], -1) if C1 else x0
    ref = F.group_norm(xc.permute(0, 2, 1), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1)
    assert rel(yn.float(), ref) < 4e-3
    assert rel(yr.float(), xc) < 4e-3
    # statistics themselves are fp32-exact: compare against the bf16 rounding of the reference
    assert (yn.float() - ref.to(torch.bfloat16).float()).abs().max() < 0.07


@pytest.mark.parametrize("B,B1,HW,C0,C1,with_sums", [(8, 4, 4096, 320, 320, True), (8, 4, 4096, 320, 320, False), (4, 2, 9216, 320, 320, False),
                                                    (6, 1, 1024, 640, 320, True)])
def test_groupnorm_shared_second_source(ops, cuda_dev, B, B1, HW, C0, C1, with_sums):
    """`x1_batch`: the concatenated second source (a UNet skip tensor) holds B1 < B images and image b reads image b % B1 --
    against the same call on the materialised repeat, bit for bit (normalised output and the raw bf16 copy), with the
    producers' per-image sums and with the kernel's own statistics pass."""
    g = torch.Generator(device="cuda").manual_seed(B + HW)
    x0 = torch.randn(B, HW, C0, device=cuda_dev, generator=g) * 2 + 0.3
    x1 = torch.randn(B1, HW, C1, device=cuda_dev, generator=g) - 0.5
    x1r = x1.repeat(B // B1, 1, 1).contiguous()
    gamma, beta = torch.randn(C0 + C1, device=cuda_dev, generator=g), torch.randn(C0 + C1, device=cuda_dev, generator=g)

    def sums(t):
        gran = 10
        f = t.double().view(t.shape[0], HW, t.shape[2] // gran, gran)
        return torch.stack([(f.sum((1, 3)) * 2.0 ** 32).round().long(), ((f * f).sum((1, 3)) * 2.0 ** 24).round().long()], -1).contiguous()
    kw = dict(groups=32, eps=1e-5, silu=True, want_raw=True)
    if with_sums:
        a, ar = ops.groupnorm(x0, gamma, beta, x1=x1, x0_stats=sums(x0), x1_stats=sums(x1), **kw)
        b, br = ops.groupnorm(x0, gamma, beta, x1=x1r, x0_stats=sums(x0), x1_stats=sums(x1r), **kw)
    else:
        a, ar = ops.groupnorm(x0, gamma, beta, x1=x1, **kw)
        b, br = ops.groupnorm(x0, gamma, beta, x1=x1r, **kw)
    assert torch.equal(a, b) and torch.equal(ar, br)
    ref = F.silu(F.group_norm(torch.cat([x0, x1r], -1).permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1)
    assert rel(a, ref) < 4e-3


@pytest.mark.parametrize("rows,C", [(8192, 320), (2048, 640), (512, 1280), (77, 1024)])
def test_layernorm(ops, cuda_dev, rows, C):
    g = torch.Generator(device="cuda").manual_seed(C)
    x = torch.randn(rows, C, device=cuda_dev, generator=g) * 3 + 1
    gamma = 1 + 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    beta = 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    y = ops.layernorm(x, gamma, beta)
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    assert rel(y.float(), ref) < 4e-3
    assert (y.float() - ref.to(torch.bfloat16).float()).abs().max() < 0.04


def test_softmax_rows(ops, cuda_dev):
    s = torch.randn(1024, 4096, device=cuda_dev) * 20
    p = ops.softmax_rows(s, 1 / math.sqrt(512))
    ref = torch.softmax(s / math.sqrt(512), -1)
    assert rel(p.float(), ref) < 4e-3


# ------------------------------------------------------------------ small pieces
def test_time_embed(ops, cuda_dev):
    g = torch.Generator(device="cuda").manual_seed(3)
    B = 5
    t = torch.tensor([958.0, 925.0, 1.0, 0.0, 496.0], device=cuda_dev)
    w1 = torch.randn(1280, 320, device=cuda_dev, generator=g) / math.sqrt(320)
    b1 = 0.02 * torch.randn(1280, device=cuda_dev, generator=g)
    w2 = torch.randn(1280, 1280, device=cuda_dev, generator=g) / math.sqrt(1280)
    b2 = 0.02 * torch.randn(1280, device=cuda_dev, generator=g)
    wa = torch.randn(2000, 1280, device=cuda_dev, generator=g) / math.sqrt(1280)
    ba = 0.02 * torch.randn(2000, device=cuda_dev, generator=g)
    out = ops.time_embed(t, w1, b1, w2, b2, wa, ba)
    k = torch.arange(160, device=cuda_dev, dtype=torch.float32)
    f = torch.exp(-math.log(10000.0) * k / 160)
    a = t[:, None] * f[None]
    sin = torch.cat([a.cos(), a.sin()], -1)
    emb = F.linear(F.silu(F.linear(sin, w1, b1)), w2, b2)
    ref = F.linear(F.silu(emb), wa, ba)
    assert rel(out, ref) < 2e-4, rel(out, ref)


@pytest.mark.parametrize("nchw", [True, False])
def test_conv_small_cin(ops, cuda_dev, nchw):
    g = torch.Generator(device="cuda").manual_seed(11)
    B, H, W, Cout = 2, 64, 64, 320
    x = torch.randn(B, 4, H, W, device=cuda_dev, generator=g)
    w = torch.randn(Cout, 4, 3, 3, device=cuda_dev, generator=g) / 6
    bias = torch.randn(Cout, device=cuda_dev, generator=g)
    xin = x if nchw else x.permute(0, 2, 3, 1).contiguous()
    o32, _ = ops.conv3x3_small_cin(xin, w.permute(0, 2, 3, 1).contiguous(), bias, nchw=nchw)
    ref = F.conv2d(x, w, bias, padding=1).permute(0, 2, 3, 1)
    assert rel(o32, ref) < 1e-5


@pytest.mark.parametrize("Cin,Cout,post", [(320, 4, False), (128, 3, True)])
def test_conv_small_cout(ops, cuda_dev, Cin, Cout, post):
    g = torch.Generator(device="cuda").manual_seed(12)
    B, H, W = 2, 64, 64
    x = rb(torch.randn(B, H, W, Cin, device=cuda_dev, generator=g))
    w = torch.randn(Cout, Cin, 3, 3, device=cuda_dev, generator=g) / math.sqrt(9 * Cin)
    bias = torch.randn(Cout, device=cuda_dev, generator=g)
    out = ops.conv3x3_small_cout(x, w.permute(0, 2, 3, 1).contiguous(), bias, postprocess=post)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, bias, padding=1)
    if post:
        ref = (ref * 0.5 + 0.5).clamp(0, 1).permute(0, 2, 3, 1)
    assert rel(out, ref) < 1e-5


def test_upsample_cast_latent_prep(ops, cuda_dev):
    x = torch.randn(2, 8, 8, 64, device=cuda_dev)
    up = ops.upsample2x(x)
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up, ref.to(torch.bfloat16))
    assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))
    z = torch.randn(2, 4, 16, 16, device=cuda_dev)
    w = torch.randn(4, 4, device=cuda_dev)
    b = torch.randn(4, device=cuda_dev)
    o = ops.vae_latent_prep(z, w, b, 1 / 0.18215)
    ref = F.conv2d(z / 0.18215, w[:, :, None, None], b).permute(0, 2, 3, 1)
    assert rel(o, ref) < 1e-5


@pytest.mark.parametrize("use_cfg,vpred", [(True, False), (False, False), (True, True)])
def test_cfg_ddpm_step_vs_oracle(ops, cuda_dev, use_cfg, vpred):
    """K7 vs the restated DDPMScheduler.step + CFG combine (oracle/sd21.py)."""
    from oracle.sd21 import DDPMSchedulerRef
    sch = DDPMSchedulerRef(prediction_type="v_prediction" if vpred else "epsilon")
    sch.set_timesteps(30)
    n = 3
    g = torch.Generator().manual_seed(0)
    for t in (958, 496, 1):
        eps2 = torch.randn(2 * n if use_cfg else n, 4, 64, 64, generator=g)
        x = torch.randn(n, 4, 64, 64, generator=g)
        noise = torch.randn(n, 4, 64, 64, generator=g)
        e = eps2[:n] + 5.0 * (eps2[n:] - eps2[:n]) if use_cfg else eps2
        ref_prev, ref_x0 = sch.step(e.double(), t, x.double(), noise.double())
        coef = torch.tensor(sch.coefficients(t), dtype=torch.float32, device=cuda_dev)
        x0 = torch.empty(n, 4, 64, 64, device=cuda_dev)
        prev = ops.cfg_ddpm_step(eps2.to(cuda_dev), x.to(cuda_dev), noise.to(cuda_dev), coef, guidance_scale=5.0,
                                 use_cfg=use_cfg, v_prediction=vpred, x0_out=x0)
        assert rel(prev.cpu(), ref_prev) < 2e-6
        assert rel(x0.cpu(), ref_x0) < 2e-6


@pytest.mark.parametrize("B,H,W,C0,C1,N", [(2, 64, 64, 320, 0, 320), (3, 8, 8, 1280, 1280, 1280), (2, 32, 32, 640, 320, 640),
                                           (1, 256, 256, 128, 0, 128),
                                           # rasters of <= 256 pixels: statistics finalized inside the apply kernel
                                           (2, 16, 16, 1280, 640, 1280), (8, 16, 16, 1280, 0, 1280), (2, 8, 8, 1280, 0, 1280)])
def test_groupnorm_from_epilogue_statistics(ops, cuda_dev, B, H, W, C0, C1, N):
    """GroupNorm fed by the row-block channel sums that the producing conv GEMM wrote in its epilogue
    (no statistics pass over the tensor) == GroupNorm that reads the tensor itself."""
    g = torch.Generator(device="cuda").manual_seed(H + C0)
    gran = math.gcd(C0, C1 or C0) // 32      # divides the group size of the (concatenated) GroupNorm and the concat offset

    def produce(cin, cout):
        x = rb(torch.randn(B, H, W, cin, device=cuda_dev, generator=g))
        w = rb(torch.randn(cout, cin, 3, 3, device=cuda_dev, generator=g) / math.sqrt(9 * cin))
        bias = torch.randn(cout, device=cuda_dev, generator=g)
        o, _, st = ops.gemm_conv(x, pack_conv_w(w), mode=ops.A_3X3, bias=bias, want_f32=True, want_stats=True,
                                 k_splits=0, workspace=torch.empty(16 << 20, dtype=torch.float32, device=cuda_dev),
                                 stats_gran=gran)
        return o.view(B, H * W, cout), st
    x0, s0 = produce(64, C0)
    x1, s1 = produce(64, C1) if C1 else (None, None)
    # the sums themselves: fixed-point per-image channel sums (int64, accumulated inside the GEMM), or row-block sums where
    # an image has more than 16384 pixels
    if s0.dtype == torch.int64:
        assert tuple(s0.shape) == (B, C0 // gran, 2)
        xg = x0.double().view(B, H * W, C0 // gran, gran)
        assert rel(fixed_sums(s0)[0], xg.sum((1, 3))) < 1e-6 and rel(fixed_sums(s0)[1], (xg ** 2).sum((1, 3))) < 1e-6
    else:
        assert H * W > 16384
        ref_s = x0.view(B * H * W // 32, 32, C0).sum(1)
        assert rel(s0[..., 0], ref_s) < 1e-5
    C = C0 + C1
    gamma = 1 + 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    beta = 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    y_fused, _ = ops.groupnorm(x0, gamma, beta, groups=32, eps=1e-5, silu=True, x1=x1, x0_stats=s0, x1_stats=s1)
    y_plain, _ = ops.groupnorm(x0, gamma, beta, groups=32, eps=1e-5, silu=True, x1=x1)
    xc = torch.cat([x0, x1], -1) if C1 else x0
    ref = F.silu(F.group_norm(xc.permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1)
    assert rel(y_fused.float(), ref) < 4e-3
    assert (y_fused.float() - y_plain.float()).abs().max() < 0.07
