"""SURVEY 8(f)-4, the LoRA-only training backward (`/root/reference/train_ID-Booth.py:1040-1075,1140-1146`; adapters
`:672-678`): every backward kernel against torch autograd of the same op in fp32, then the adapter gradients of the
reference's denoising loss through the whole UNet against the oracle (`oracle/lora_grad.py`: autograd through the
restated UNet, itself anchored by fp64 finite differences in tests/test_structure_cpu.py).

Tolerances: operands of the backward GEMMs are bf16 like the forward's, so op-level gradients agree to ~1e-2 relative;
for the end-to-end adapter gradients the gate is per-adapter cosine similarity and relative L2 (stated in the test)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rb(t):
    return t.to(torch.bfloat16)


@pytest.fixture(scope="module")
def ops(cuda_dev):
    from faceposegenerator_b200 import ops as o
    return o


@pytest.mark.parametrize("B,heads,Tq,Tkv", [(2, 5, 512, 512), (1, 10, 1024, 1024), (2, 3, 256, 77), (1, 2, 200, 333), (3, 20, 64, 64),
                                            (2, 5, 4096, 77), (1, 5, 1024, 1024)])
def test_attention_backward(ops, cuda_dev, B, heads, Tq, Tkv):
    C = heads * 64
    g = torch.Generator(device="cuda").manual_seed(Tq + Tkv)
    self_attn = Tq == Tkv
    if self_attn:
        qkv = rb(torch.randn(B * Tq, 3 * C, device=cuda_dev, generator=g))
        q_t, k_t, v_t, cq, ck, cv = qkv, qkv, qkv, 0, C, 2 * C
        q, k, v = (x.clone().requires_grad_(True) for x in qkv.float().view(B, Tq, 3, heads, 64).unbind(2))
    else:
        qq = rb(torch.randn(B * Tq, C, device=cuda_dev, generator=g))
        kv = rb(torch.randn(B * Tkv, 2 * C, device=cuda_dev, generator=g))
        q_t, k_t, v_t, cq, ck, cv = qq, kv, kv, 0, 0, C
        q = qq.float().view(B, Tq, heads, 64).clone().requires_grad_(True)
        k, v = (x.clone().requires_grad_(True) for x in kv.float().view(B, Tkv, 2, heads, 64).unbind(2))
    d_o = rb(torch.randn(B * Tq, C, device=cuda_dev, generator=g))
    lse = torch.empty(B, heads, Tq, device=cuda_dev)
    o = ops.attention(q_t, k_t, v_t, batch=B, heads=heads, t_q=Tq, t_kv=Tkv, scale=0.125, col0_q=cq, col0_k=ck, col0_v=cv, lse=lse)
    # reference: explicit softmax in fp32 (+ its log-sum-exp in the log2 domain)
    s = torch.einsum("bqhd,bkhd->bhqk", q, k) * 0.125
    ref_lse = torch.logsumexp(s, -1) / math.log(2.0)
    ref = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, -1), v).reshape(B * Tq, C)
    assert rel(o.float(), ref) < 6e-3
    assert (lse - ref_lse).abs().max() < 2e-2
    ref.backward(d_o.float())
    dq, dk, dv = ops.attention_backward(q_t, k_t, v_t, o, d_o, lse, batch=B, heads=heads, t_q=Tq, t_kv=Tkv, scale=0.125,
                                        col0_q=cq, col0_k=ck, col0_v=cv)
    e = (rel(dq, q.grad.reshape(B * Tq, C)), rel(dk.float(), k.grad.reshape(B * Tkv, C)), rel(dv.float(), v.grad.reshape(B * Tkv, C)))
    print(f"attention backward B{B} h{heads} Tq{Tq} Tkv{Tkv}: dq {e[0]:.2e} dk {e[1]:.2e} dv {e[2]:.2e}")
    assert max(e) < 1.5e-2


@pytest.mark.parametrize("rows,C", [(4096, 320), (1000, 640), (77, 1280)])
def test_layernorm_backward(ops, cuda_dev, rows, C):
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = (torch.randn(rows, C, device=cuda_dev, generator=g) * 2 + 0.3).requires_grad_(True)
    gamma = 1 + 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    beta = 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    dy = torch.randn(rows, C, device=cuda_dev, generator=g)
    F.layer_norm(x, (C,), gamma, beta, 1e-5).backward(dy)
    base = torch.randn(rows, C, device=cuda_dev, generator=g)
    dx = ops.layernorm_backward(dy, x.detach(), gamma)
    acc = ops.layernorm_backward(dy, x.detach(), gamma, dx=base.clone(), add=True)
    assert rel(dx, x.grad) < 1e-5
    assert rel(acc, x.grad + base) < 1e-5


@pytest.mark.parametrize("B,HW,C0,C1,silu,eps", [(2, 1024, 320, 0, True, 1e-5), (2, 256, 640, 320, True, 1e-5), (1, 64, 1280, 1280, True, 1e-5),
                                                (2, 1024, 320, 0, False, 1e-6)])
def test_groupnorm_backward(ops, cuda_dev, B, HW, C0, C1, silu, eps):
    g = torch.Generator(device="cuda").manual_seed(HW + C0)
    x0 = (torch.randn(B, HW, C0, device=cuda_dev, generator=g) * 2 + 0.5).requires_grad_(True)
    x1 = torch.randn(B, HW, C1, device=cuda_dev, generator=g).requires_grad_(True) if C1 else None
    C = C0 + C1
    gamma = 1 + 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    beta = 0.1 * torch.randn(C, device=cuda_dev, generator=g)
    dy = torch.randn(B, HW, C, device=cuda_dev, generator=g)
    xc = torch.cat([x0, x1], -1) if C1 else x0
    y = F.group_norm(xc.permute(0, 2, 1), 32, gamma, beta, eps)
    y = (F.silu(y) if silu else y).permute(0, 2, 1)
    y.backward(dy)
    xg = xc.detach().double().view(B, HW, 32, C // 32)
    mean = xg.mean((1, 3))
    rstd = 1.0 / torch.sqrt(xg.var((1, 3), unbiased=False) + eps)
    stats = torch.stack([mean, rstd], -1).float().contiguous()
    dx0, dx1 = ops.groupnorm_backward(dy, x0.detach(), gamma, beta, stats, groups=32, silu=silu, x1=None if x1 is None else x1.detach())
    assert rel(dx0, x0.grad) < 2e-5
    if C1:
        assert rel(dx1, x1.grad) < 2e-5
    # statistics from the fixed-point sums a producing GEMM would have accumulated
    gran = math.gcd(C0, C1 or C0) // 32

    def fx(x):
        d = x.detach().double().view(B, HW, -1, gran)
        return torch.stack([(d.sum((1, 3)) * 2.0 ** 32).round().long(), ((d * d).sum((1, 3)) * 2.0 ** 24).round().long()], -1).contiguous()
    st2 = ops.group_stats_from_sums(fx(x0), C0, HW, 32, eps, fx(x1) if C1 else None, C1)
    assert rel(st2, stats) < 1e-5


def test_geglu_backward(ops, cuda_dev):
    M, H = 512, 1280
    g = torch.Generator(device="cuda").manual_seed(1)
    a = rb(torch.randn(M, H, device=cuda_dev, generator=g)).float().requires_grad_(True)
    gt = rb(torch.randn(M, H, device=cuda_dev, generator=g)).float().requires_grad_(True)
    dh = rb(torch.randn(M, H, device=cuda_dev, generator=g))
    (a * F.gelu(gt)).backward(dh.float())
    # interleave [a(16) | g(16)] blocks like the fused FF-in GEMM lays its output out
    u = torch.stack([a.detach().view(M, H // 16, 16), gt.detach().view(M, H // 16, 16)], 2).reshape(M, 2 * H).to(torch.bfloat16)
    du = ops.geglu_backward(dh, u).float().view(M, H // 16, 2, 16)
    assert rel(du[:, :, 0].reshape(M, H), a.grad) < 6e-3
    assert rel(du[:, :, 1].reshape(M, H), gt.grad) < 6e-3


@pytest.mark.parametrize("M,W,r", [(4096, 320, 4), (1000, 1280, 4), (154, 1024, 8)])
def test_lora_wgrad(ops, cuda_dev, M, W, r):
    g = torch.Generator(device="cuda").manual_seed(M)
    wide = rb(torch.randn(M, W + 64, device=cuda_dev, generator=g))
    skinny = rb(torch.randn(M, 32, device=cuda_dev, generator=g))
    ref = wide[:, 64:].float().t() @ skinny[:, :r].float()
    out = ops.lora_wgrad(wide, skinny, r, col0_w=64, width=W, scale=0.5)
    assert rel(out, 0.5 * ref) < 1e-5
    out_t = ops.lora_wgrad(wide, skinny, r, col0_w=64, width=W, transpose_out=True)
    assert rel(out_t, ref.t()) < 1e-5
    again = ops.lora_wgrad(wide, skinny, r, col0_w=64, width=W, scale=0.5)
    assert torch.equal(out, again)        # fixed summation order


def test_conv_input_gradients(ops, cuda_dev):
    """Input gradients of the three convolution kinds on `idb_gemm_conv`: stride-1 (flipped / transposed weight), stride-2
    (zero-inserted gradient) and nearest-2x-upsample + conv (2x2 sum-pool of the full-resolution gradient)."""
    from faceposegenerator_b200.lora_backward import pack_conv_dgrad_weight
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W, Cin, Cout = 2, 16, 16, 128, 64
    w = rb(torch.randn(Cout, Cin, 3, 3, device=cuda_dev, generator=g) / math.sqrt(9 * Cin)).float()
    wd = pack_conv_dgrad_weight(w, cuda_dev)
    x = torch.randn(B, Cin, H, W, device=cuda_dev, generator=g, requires_grad=True)
    # stride 1
    dy = rb(torch.randn(B, H, W, Cout, device=cuda_dev, generator=g))
    F.conv2d(x, w, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    dx = ops.gemm_conv(dy, wd, mode=ops.A_3X3, want_f32=True)[0].view(B, H, W, Cin)
    assert rel(dx.permute(0, 3, 1, 2), x.grad) < 4e-3
    # stride 2
    x.grad = None
    dy2 = rb(torch.randn(B, H // 2, W // 2, Cout, device=cuda_dev, generator=g))
    F.conv2d(x, w, stride=2, padding=1).backward(dy2.float().permute(0, 3, 1, 2))
    dx2 = ops.gemm_conv(ops.zero_insert2x(dy2), wd, mode=ops.A_3X3, want_f32=True)[0].view(B, H, W, Cin)
    assert rel(dx2.permute(0, 3, 1, 2), x.grad) < 4e-3
    # nearest 2x + conv
    x.grad = None
    dy3 = rb(torch.randn(B, 2 * H, 2 * W, Cout, device=cuda_dev, generator=g))
    F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1).backward(dy3.float().permute(0, 3, 1, 2))
    d_up = ops.gemm_conv(dy3, wd, mode=ops.A_3X3, want_f32=True)[0].view(B, 2 * H, 2 * W, Cin)
    dx3 = ops.sumpool2x(d_up)
    assert rel(dx3.permute(0, 3, 1, 2), x.grad) < 4e-3


def test_lora_linear_input_gradient(ops, cuda_dev):
    """dX = dY W + s (dY B) A of an adapted fused q/k/v projection through the fused-LoRA GEMM with swapped adapter roles."""
    from faceposegenerator_b200.lora_backward import pack_lora_dgrad
    g = torch.Generator(device="cuda").manual_seed(4)
    M, C, r = 2048, 320, 4
    W = rb(torch.randn(3 * C, C, device=cuda_dev, generator=g) / math.sqrt(C)).float()
    ads = [(rb(torch.randn(r, C, generator=g, device=cuda_dev) * 0.25).float(), rb(torch.randn(C, r, generator=g, device=cuda_dev) * 0.1).float(), 0.5 + k)
           for k in range(3)]
    dy = rb(torch.randn(M, 3 * C, device=cuda_dev, generator=g))
    ld, lu = pack_lora_dgrad([(a.cpu(), b.cpu(), s) for a, b, s in ads], C, C, cuda_dev)
    dx = ops.gemm_conv(dy, W.t().contiguous().to(torch.bfloat16), lora_down=ld, lora_up=lu, lora_seg_n=C, want_f32=True)[0]
    ref = dy.float() @ W
    for k, (a, b, s) in enumerate(ads):
        ref = ref + s * (dy.float()[:, k * C:(k + 1) * C] @ b) @ a
    assert rel(dx, ref) < 4e-3


def test_unet_lora_gradients_vs_oracle(cuda_dev):
    """The adapter gradients of the reference's denoising loss (epsilon target, `F.mse_loss(..., "mean")`,
    train_ID-Booth.py:1055-1075) for a batch of two 64 x 64 latents with different timesteps: hand-written backward vs
    autograd through the fp32 oracle UNet on the GPU.  Gate: the prediction itself within the forward tolerance; per
    adapter cosine >= 0.98 for both factors; relative L2 of the concatenated gradient <= 5e-2 (bf16 operands in ~200
    chained backward GEMMs; the gradient noise of bf16 training is of this order)."""
    from faceposegenerator_b200.lora_backward import UNetLoRAGrad
    from faceposegenerator_b200.unet import UNet2DConditionModel
    from faceposegenerator_b200.weights import random_lora, random_state_dict, unet_manifest
    from oracle import lora_grad, sd21
    sd = random_state_dict(unet_manifest(), 0)
    lora = random_lora(seed=1)
    unet = UNet2DConditionModel(sd, device=cuda_dev)
    unet.set_lora(lora)
    g = torch.Generator().manual_seed(21)
    x0 = torch.randn(2, 4, 64, 64, generator=g)
    noise = torch.randn(2, 4, 64, 64, generator=g)
    ctx = torch.randn(2, 77, 1024, generator=g)
    t = torch.tensor([620, 85])
    sch = sd21.DDPMSchedulerRef()
    noisy = sch.add_noise(x0, noise, t)
    sd_gpu = {k: v.to(cuda_dev) for k, v in sd.items()}
    lora_gpu = {k: (d.to(cuda_dev), u.to(cuda_dev), s) for k, (d, u, s) in lora.items()}
    loss_ref, gref = lora_grad.lora_gradients(sd_gpu, lora_gpu, x0.to(cuda_dev), noise.to(cuda_dev), t.to(cuda_dev), ctx.to(cuda_dev),
                                              sd21.UNET_SD21)
    eng = UNetLoRAGrad(unet, lora)
    eps = eng.forward(noisy.to(cuda_dev), t.float().to(cuda_dev), ctx.to(cuda_dev))
    with torch.no_grad():
        ref_eps = torch.cat([sd21.unet_forward(sd_gpu, noisy[i:i + 1].to(cuda_dev), int(t[i]), ctx[i:i + 1].to(cuda_dev), lora_gpu) for i in range(2)])
    assert rel(eps, ref_eps) < 1.2e-2       # sanity only: the forward's own 1e-2 gate is tests/test_model_gpu.py (this tape forward takes the plain upsample path)
    loss = F.mse_loss(eps, noise.to(cuda_dev))
    d_eps = 2.0 * (eps - noise.to(cuda_dev)) / eps.numel()
    grads = eng.backward(d_eps)
    assert set(grads) == set(lora)
    print(f"loss {float(loss):.6f} (oracle {float(loss_ref):.6f})")
    worst_cos, cat_a, cat_b = 1.0, [], []
    for k in sorted(lora):
        dA, dB = grads[k]
        rA, rB = gref[k]
        for mine, theirs in ((dA, rA), (dB, rB)):
            cos = float(F.cosine_similarity(mine.flatten().double().cpu(), theirs.flatten().double().cpu(), dim=0))
            worst_cos = min(worst_cos, cos)
            cat_a.append(mine.flatten().double().cpu())
            cat_b.append(theirs.flatten().double().cpu())
    e = rel(torch.cat(cat_a), torch.cat(cat_b))
    print(f"LoRA gradients: worst per-tensor cosine {worst_cos:.4f}, relative L2 of all {len(cat_a)} tensors {e:.3e}")
    assert abs(float(loss) - float(loss_ref)) < 1e-2 * float(loss_ref)
    assert worst_cos >= 0.98
    assert e <= 5e-2


def test_lora_trainer_reduces_the_denoising_loss(cuda_dev):
    """`LoRATrainer.step` = add_noise -> UNet -> MSE -> backward -> clip -> AdamW (train_ID-Booth.py:1012-1146) on a fixed
    batch: the loss must go down, the adapters must move, the frozen base weights must not."""
    from faceposegenerator_b200 import DDPMScheduler
    from faceposegenerator_b200.lora_backward import LoRATrainer
    from faceposegenerator_b200.unet import UNet2DConditionModel
    from faceposegenerator_b200.weights import random_lora, random_state_dict, unet_manifest
    unet = UNet2DConditionModel(random_state_dict(unet_manifest(), 0), device=cuda_dev)
    lora = random_lora(seed=1)
    base = unet.transformers[0].w_qkv.clone()
    tr = LoRATrainer(unet, lora, DDPMScheduler.from_pretrained("stabilityai/stable-diffusion-2-1-base", subfolder="scheduler"), lr=1e-3)
    g = torch.Generator().manual_seed(2)
    x0, noise = torch.randn(2, 4, 64, 64, generator=g), torch.randn(2, 4, 64, 64, generator=g)
    ctx, t = torch.randn(2, 77, 1024, generator=g), torch.tensor([500, 120])
    losses = [tr.step(x0, noise, t, ctx.to(cuda_dev))[0] for _ in range(6)]
    print("denoising loss over 6 AdamW steps:", [round(v, 5) for v in losses])
    assert losses[-1] < losses[0] and min(losses[1:]) < losses[0]
    moved = max(float((tr.params[k][1].detach().cpu() - lora[k][1]).abs().max()) for k in lora)
    assert moved > 1e-4
    assert torch.equal(unet.transformers[0].w_qkv, base)


def test_lora_trainer_graph_replay_equals_eager_steps(cuda_dev):
    """The trainer captures add_noise -> forward with tape -> loss -> backward into one CUDA graph per batch geometry and
    re-installs the adapters in place after every AdamW step, so the graph is replayed with NEW timesteps / data each
    step.  Its losses and gradient norms must follow the eager trainer's (same start, same batches; the only
    non-bit-reproducible piece is the fp32 atomic accumulation of dQ in the attention backward)."""
    from faceposegenerator_b200 import DDPMScheduler
    from faceposegenerator_b200.lora_backward import LoRATrainer
    from faceposegenerator_b200.unet import UNet2DConditionModel
    from faceposegenerator_b200.weights import random_lora, random_state_dict, unet_manifest
    sched = DDPMScheduler.from_pretrained("stabilityai/stable-diffusion-2-1-base", subfolder="scheduler")
    g = torch.Generator().manual_seed(9)
    batches = [(torch.randn(2, 4, 64, 64, generator=g), torch.randn(2, 4, 64, 64, generator=g),
                torch.randint(0, 1000, (2,), generator=g), torch.randn(2, 77, 1024, generator=g)) for _ in range(4)]
    runs = {}
    for graph in (True, False):
        unet = UNet2DConditionModel(random_state_dict(unet_manifest(), 0), device=cuda_dev)
        tr = LoRATrainer(unet, random_lora(seed=1), sched, lr=1e-3, use_cuda_graph=graph)
        runs[graph] = [tr.step(x0, n, t, c.to(cuda_dev)) for (x0, n, t, c) in batches]
        if graph:
            assert len(tr._graphs) == 1, "one capture for the four steps"
    for (lg, ng), (le, ne) in zip(runs[True], runs[False]):
        assert abs(lg - le) <= 2e-3 * abs(le) and abs(ng - ne) <= 2e-2 * abs(ne), (runs[True], runs[False])
