"""T0 (CPU, no GPU): oracle vs known answers / golden vectors, host logic, C-ABI symbol check."""
import math
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TIMESTEPS_30 = [958, 925, 892, 859, 826, 793, 760, 727, 694, 661, 628, 595, 562, 529, 496, 463, 430, 397, 364, 331,
                298, 265, 232, 199, 166, 133, 100, 67, 34, 1]
# (t -> sqrt_acp, sqrt_1m_acp, c_x0, c_xt, sigma)
COEF = {958: (0.08680304, 0.99622548, 0.03215177, 0.83015662, 0.55242407),
        925: (0.10421189, 0.99455512, 0.03676465, 0.83679783, 0.54037559),
        34: (0.98380959, 0.17921701, 0.94775540, 0.05223803, 0.04020363),
        1: (0.99914765, 0.04127926, 1.0, 0.0, 1.0e-10)}


def test_parameter_counts():
    from faceposegenerator_b200 import weights as w
    assert sum(math.prod(s) for _, s in w.unet_manifest()) == 865_910_724
    assert sum(math.prod(s) for _, s in w.vae_decoder_manifest()) == 49_490_199
    assert sum(math.prod(s) for _, s in w.vae_encoder_manifest()) == 34_163_592 + 72   # decoder + encoder = 83,653,863
    lm = w.lora_manifest()
    assert len(lm) == 128 and sum(4 * (a + b) for _, a, b in lm) == 829_952
    from faceposegenerator_b200.text import text_manifest
    assert sum(math.prod(s) for _, s in text_manifest()) == 340_387_840
    names = [n for n, _ in w.unet_manifest()]
    assert len(names) == len(set(names))
    assert sum(1 for n in names if re.search(r"resnets\.\d+\.conv1\.weight$", n)) == 22
    assert sum(1 for n in names if n.endswith("proj_in.weight")) == 16


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_scheduler_known_answers(which):
    if which == "oracle":
        from oracle.sd21 import DDPMSchedulerRef
        s = DDPMSchedulerRef()
        s.set_timesteps(30)
    else:
        from faceposegenerator_b200 import DDPMScheduler
        s = DDPMScheduler.from_pretrained("stabilityai/stable-diffusion-2-1-base", subfolder="scheduler")
        s.set_timesteps(30)
    assert s.timesteps.tolist() == TIMESTEPS_30
    assert abs(float(s.betas[0]) - 0.00085) < 1e-9 and abs(float(s.betas[999]) - 0.012) < 1e-8
    for t, a in ((0, 0.99914998), (1, 0.99829602), (34, 0.96788126), (925, 0.01086012), (958, 0.00753477), (999, 0.00466010)):
        assert abs(float(s.alphas_cumprod[t]) - a) < 2e-7
    for t, ref in COEF.items():
        got = s.coefficients(t)
        for g, r in zip(got, ref):
            assert abs(g - r) < 2e-7 * max(1.0, abs(r)) + 1e-12, (t, got, ref)
    assert s.previous_timestep(958) == 925 and s.previous_timestep(1) == -1
    s2 = type(s)()
    assert s2.previous_timestep(500) == 499   # training use: no set_timesteps (train_ID-Booth.py:1081)


def test_sinusoid_known_answers():
    from oracle.sd21 import timestep_sinusoid
    e = timestep_sinusoid(torch.tensor([958.0]))[0]
    ref_cos = [-0.98279631, 0.93290263, 0.76819366, -0.23576230]
    ref_sin = [0.18469287, -0.36012876, -0.64021748, 0.97181076]
    assert torch.allclose(e[:4], torch.tensor(ref_cos), atol=2e-4)
    assert torch.allclose(e[160:164], torch.tensor(ref_sin), atol=2e-4)


def test_oracle_scheduler_step_matches_closed_form():
    from oracle.sd21 import DDPMSchedulerRef
    s = DDPMSchedulerRef()
    s.set_timesteps(30)
    g = torch.Generator().manual_seed(0)
    x, eps, z = (torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64) for _ in range(3))
    prev, x0 = s.step(eps, 958, x, z)
    sa, sb, c0, ct, sg = COEF[958]
    x0_ref = (x - sb * eps) / sa
    assert torch.allclose(x0, x0_ref, rtol=1e-6)
    assert torch.allclose(prev, c0 * x0_ref + ct * x + sg * z, rtol=1e-5, atol=1e-6)
    # add_noise / training-mode previous timestep
    n = s.add_noise(x.float(), z.float(), torch.tensor([10, 900]))
    a = s.alphas_cumprod[torch.tensor([10, 900])].view(2, 1, 1, 1)
    assert torch.allclose(n, a.sqrt() * x.float() + (1 - a).sqrt() * z.float(), atol=1e-6)


TINY = dict(in_channels=4, out_channels=4, block_out_channels=(32, 64, 64, 64), down_attn=(True, True, True, False),
            up_attn=(False, True, True, True), layers_per_block=2, head_dim=8, cross_attention_dim=48,
            norm_num_groups=8, norm_eps=1e-5, time_embed_in=32)
TINY_M = dict(sample_size=8, in_channels=4, out_channels=4, block_out_channels=(32, 64, 64, 64),
              down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
              up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3, layers_per_block=2,
              cross_attention_dim=48, norm_num_groups=8, norm_eps=1e-5)


def _tiny():
    from faceposegenerator_b200 import weights as w
    sd = {k: v.double() for k, v in w.random_state_dict(w.unet_manifest(TINY_M), 3).items()}
    lora = {k: (d.double(), (u * 20).double(), s) for k, (d, u, s) in w.random_lora(TINY_M, seed=3).items()}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 16, 16, generator=g, dtype=torch.float64)
    ctx = torch.randn(2, 5, 48, generator=g, dtype=torch.float64)
    return sd, lora, x, ctx


def test_oracle_lora_metamorphic_fp64():
    """merged == unmerged (peft semantics); B == 0 is an exact no-op; rows are independent."""
    from oracle import sd21
    sd, lora, x, ctx = _tiny()
    with torch.no_grad():
        base = sd21.unet_forward(sd, x, 500, ctx, None, TINY)
        un = sd21.unet_forward(sd, x, 500, ctx, lora, TINY)
        me = sd21.unet_forward(sd21.merge_lora(sd, lora), x, 500, ctx, None, TINY)
        zero = {k: (d, torch.zeros_like(u), s) for k, (d, u, s) in lora.items()}
        z = sd21.unet_forward(sd, x, 500, ctx, zero, TINY)
        row0 = sd21.unet_forward(sd, x[:1], 500, ctx[:1], lora, TINY)
    assert (un - me).abs().max() < 1e-10
    assert (un - base).abs().max() > 1e-4
    assert torch.equal(z, base)
    assert (row0 - un[:1]).abs().max() < 1e-10


def test_oracle_lora_gradients_fp64_finite_differences():
    """Checker for SURVEY 8(f)-4 (LoRA-only backward, not built yet): autograd gradients of the reference's denoising loss
    with respect to the adapters (oracle/lora_grad.py) against central finite differences in fp64, and the structure of
    the up-projection gradient (d loss / d B = sum dY (x A^T): zero wherever x A^T is zero)."""
    from oracle import lora_grad
    sd, lora, x, ctx = _tiny()
    g = torch.Generator().manual_seed(5)
    noise = torch.randn(x.shape, generator=g, dtype=torch.float64)
    t = torch.tensor([700, 40])
    loss, grads = lora_grad.lora_gradients(sd, lora, x, noise, t, ctx, TINY)
    assert set(grads) == set(lora) and float(loss) > 0
    keys = sorted(lora)
    picks = [keys[0], keys[len(keys) // 2], keys[-1]]          # adapters in different blocks / on different projections
    h = 1e-6
    for k in picks:
        for which, idx in ((0, (1, 3)), (1, (2, 1))):
            def loss_at(delta):
                d, u, s = lora[k]
                d, u = d.clone(), u.clone()
                (d if which == 0 else u)[idx] += delta
                with torch.no_grad():
                    return float(lora_grad.denoising_loss(sd, {**lora, k: (d, u, s)}, x, noise, t, ctx, TINY))
            fd = (loss_at(h) - loss_at(-h)) / (2 * h)
            an = float(grads[k][which][idx])
            assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)) + 1e-9, (k, which, fd, an)
    # an adapter whose down-projection is zero has a zero up-projection gradient (and a non-zero down gradient)
    k = picks[1]
    zeroed = {**lora, k: (torch.zeros_like(lora[k][0]), lora[k][1], lora[k][2])}
    _, gz = lora_grad.lora_gradients(sd, zeroed, x, noise, t, ctx, TINY)
    assert float(gz[k][1].abs().max()) == 0.0 and float(gz[k][0].abs().max()) > 0.0
    # v-prediction target (train_ID-Booth.py:1057-1058) gives a different loss on the same prediction
    lv = lora_grad.denoising_loss(sd, lora, x, noise, t, ctx, TINY, prediction_type="v_prediction")
    assert abs(float(lv) - float(loss)) > 1e-6


def test_oracle_cfg_scale_one_is_conditional_branch():
    from oracle import sd21
    sd, lora, _, _ = _tiny()
    sd = {k: v.float() for k, v in sd.items()}
    g = torch.Generator().manual_seed(1)
    tape = torch.randn(4, 1, 4, 16, 16, generator=g)
    pe, ne = torch.randn(1, 5, 48, generator=g), torch.randn(1, 5, 48, generator=g)
    with torch.no_grad():
        lat, eps = sd21.denoise_loop(sd, None, pe, ne, tape, num_steps=3, guidance_scale=1.0, cfg=TINY)
        cond = sd21.unet_forward(sd, tape[0], 667, pe, None, TINY)
    assert torch.allclose(eps[0], cond, atol=1e-5)
    assert lat.shape == (4, 1, 4, 16, 16)


def test_iresnet_oracle_vs_reference_golden():
    """The one PINNED oracle: outputs of /root/reference/ArcFace_files/backbones/iresnet.py itself
    (fixture made by tests/golden/make_iresnet_golden.py)."""
    from oracle.iresnet import iresnet_forward, keyed_state_dict
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "iresnet100_golden.pt"))
    # the key/shape set without importing the reference: shapes follow iresnet.py:67-162
    from faceposegenerator_b200.weights import iresnet_manifest
    shapes = dict(iresnet_manifest("r100"))
    assert len(shapes) == gold["n_state"] == 925
    sd = keyed_state_dict({k: torch.zeros(s, dtype=torch.long if k.endswith("tracked") else torch.float32)
                           for k, s in shapes.items()}, seed=0)
    assert sum(v.numel() for k, v in sd.items() if "running" not in k and "tracked" not in k) == gold["n_params"] == 65_156_160
    x = torch.randn(2, 3, 112, 112, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        y = iresnet_forward(sd, x)
    assert torch.allclose(y, gold["embedding"], rtol=1e-4, atol=1e-3 * float(gold["embedding"].abs().mean()))


def test_clip_text_oracle_vs_transformers_golden():
    """PINNED: oracle/clip_text.py against outputs of transformers' own `CLIPTextModel` with the SD2.1-base text config
    (fixture made by tests/golden/make_clip_text_golden.py; weights regenerated from key names, ids from the fixture)."""
    from faceposegenerator_b200.text import text_manifest
    from faceposegenerator_b200.weights import random_state_dict
    from oracle.clip_text import clip_text_forward
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "clip_text_golden.pt"))
    sd = random_state_dict(text_manifest(), 0)
    assert sum(v.numel() for v in sd.values()) == gold["n_params"] == 340_387_840
    y = clip_text_forward(sd, gold["ids"])
    ref = gold["last_hidden_state"]
    e = float((y - ref).norm() / ref.norm())
    assert e < 2e-5, e
    assert torch.allclose(y, ref, atol=2e-4)


def _decoded_image(seed, size):   # same seeded stand-in for `vae.decode(z).sample` as tests/golden/make_arcface_glue_golden.py
    return torch.rand(1, 3, size, size, generator=torch.Generator().manual_seed(seed)) * 2.4 - 1.2


def test_arcface_glue_oracle_vs_reference_golden():
    """PINNED: oracle/arcface_glue.py against outputs of the reference's own `latents_to_image_for_mtcnn` /
    `cropped_image_to_arcface_input` (train_ID-Booth.py:433-455) and its bbox crop (`:1090`)."""
    from oracle.arcface_glue import crop_to_arcface_input, decoded_to_mtcnn_image
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "arcface_glue_golden.pt"))
    for k, (seed, size, bbox) in enumerate(gold["cases"]):
        img = decoded_to_mtcnn_image(_decoded_image(seed, size))
        assert torch.equal(img[::16, ::16], gold["mtcnn_image_slices"][k])
        x = crop_to_arcface_input(img, bbox)[0]
        want = gold["arcface_inputs"][k]
        got = x if want.shape[-1] == 112 else x[:, ::4, ::4]
        assert torch.allclose(got, want, atol=1e-6), float((got - want).abs().max())


def test_snapshot_tokenizer_is_used_when_the_model_directory_ships_one(tmp_path):
    """A local SD2.1 snapshot brings `tokenizer/{vocab.json, merges.txt}`: prompts are then tokenized by transformers'
    CLIPTokenizer the way `encode_prompt` / train_ID-Booth.py:463-469 do (max_length padding to 77, truncation); without a
    snapshot the hashed stand-in is used."""
    import json
    from faceposegenerator_b200.text import HashTokenizer, SnapshotTokenizer, load_tokenizer
    assert isinstance(load_tokenizer(None), HashTokenizer) and isinstance(load_tokenizer(str(tmp_path)), HashTokenizer)
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    cs = [chr(b) for b in bs] + [chr(256 + k) for k in range(256 - len(bs))]     # CLIP's byte-level alphabet
    merges = [("f", "a"), ("fa", "c"), ("fac", "e</w>"), ("p", "h"), ("ph", "o"), ("pho", "t"), ("phot", "o</w>")]
    vocab = cs + [c + "</w>" for c in cs] + [a + b for a, b in merges] + ["<|startoftext|>", "<|endoftext|>"]
    tok_dir = tmp_path / "tokenizer"
    tok_dir.mkdir()
    (tok_dir / "vocab.json").write_text(json.dumps({t: i for i, t in enumerate(vocab)}))
    (tok_dir / "merges.txt").write_text("#version: 0.2\n" + "\n".join(f"{a} {b}" for a, b in merges) + "\n")
    (tok_dir / "tokenizer_config.json").write_text(json.dumps({
        "model_max_length": 77, "pad_token": "!", "bos_token": "<|startoftext|>", "eos_token": "<|endoftext|>",
        "unk_token": "<|endoftext|>", "tokenizer_class": "CLIPTokenizer"}))
    tok = load_tokenizer(str(tmp_path))
    assert isinstance(tok, SnapshotTokenizer) and tok.model_max_length == 77
    ids = tok(["face photo", "photo face " * 60])
    bos, eos, face, photo = vocab.index("<|startoftext|>"), vocab.index("<|endoftext|>"), vocab.index("face</w>"), vocab.index("photo</w>")
    assert ids.shape == (2, 77) and ids.dtype == torch.long
    assert ids[0].tolist() == [bos, face, photo, eos] + [0] * 73           # SD2.1 pads with "!" (id 0), not with EOS
    assert ids[1, 0] == bos and ids[1, -1] == eos and ids[1, 1:5].tolist() == [photo, face, photo, face]   # truncated to 77


def test_postprocess_oracle_vs_reference_golden():
    """PINNED (row a14): oracle `postprocess_np` + the uint8 conversion of `output_type="pil"` against the reference's own
    `latents_to_pil_images` (train_ID-Booth.py:408-417) run on the same seeded decoder outputs."""
    from oracle.sd21 import postprocess_np
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "arcface_glue_golden.pt"))
    batch = torch.cat([_decoded_image(10, 256), _decoded_image(11, 256)])
    img = postprocess_np(batch)                                  # float32 [2, 256, 256, 3] in [0, 1]
    assert img.shape == (2, 256, 256, 3) and img.min() >= 0.0 and img.max() <= 1.0
    u8 = torch.from_numpy((img * 255).round().astype("uint8"))
    assert torch.equal(u8[:, ::4, ::4], gold["pil_uint8_slices"])


def test_triplet_identity_loss_is_the_reference_loss_object():
    """`triplet_identity_loss` == the loss object the reference builds at train_ID-Booth.py:974-979
    (`TripletMarginWithDistanceLoss` over 1 - cosine) on the shapes of its call site (`:1133`)."""
    import torch.nn.functional as F
    from faceposegenerator_b200.iresnet import identity_loss, identity_noise_level_weight, triplet_identity_loss

    def cosine_distance(x, y):
        return 1 - F.cosine_similarity(x, y)
    ref = torch.nn.TripletMarginWithDistanceLoss(distance_function=cosine_distance)
    g = torch.Generator().manual_seed(4)
    for n in (1, 4):
        a, gt = torch.randn(n, 512, generator=g), torch.randn(2, 512, generator=g)
        assert torch.allclose(triplet_identity_loss(a, gt[0][None, :], gt[1][None, :]), ref(a, gt[0][None, :], gt[1][None, :]), atol=1e-7)
        close = a + 0.05 * torch.randn(n, 512, generator=g)       # a positive near the anchor: the hinge is active
        assert torch.allclose(triplet_identity_loss(a, close, gt[1][None, :]), ref(a, close, gt[1][None, :]), atol=1e-7)
        cos = torch.nn.CosineSimilarity(dim=1, eps=1e-6)
        assert torch.allclose(identity_loss(a, gt[0]), 1 - cos(a, gt[0][None]), atol=1e-6)      # `:1096-1098`
    assert identity_noise_level_weight(250) == (1 - 250 / 1000) ** 2 and identity_noise_level_weight(250, 1000, False) == 1


def test_epilogue_statistics_gate_mirrors_the_kernel_precondition():
    """`ops.epilogue_stats_supported` (host mirror of the `stats_partials` check in gemm_tc.cu): every raster of the 512 x 512
    path keeps the fused GroupNorm statistics; the 96 / 48 / 24 / 12-wide rasters of config 4 (768 x 768) and the
    VAE's 192-wide level fall back to GroupNorm's own statistics kernel instead of raising IDB_E_UNSUPPORTED."""
    from faceposegenerator_b200.ops import epilogue_stats_supported as ok
    for b in (1, 2, 4, 8, 16):
        assert all(ok(b, w, w) for w in (8, 16, 32, 64, 128, 256, 512, 384, 768))
        assert not any(ok(b, w, w) for w in (12, 24, 48, 96, 192, 112, 56))
    assert ok(1, 1, 8 * 4096) and ok(1, 1, 16 * 9216) and not ok(1, 1, 154)      # Linear layers: rows in whole 32-row blocks


def test_upsample_four_phase_weights_are_the_upsampled_conv():
    """Host packing of the Upsample2D rewrite: four 2x2 convolutions on the low-resolution tensor (tap offsets (a-1, c-1),
    `pack_upsample_phase_weights`) == conv3x3(pad 1) on the nearest-2x upsampled tensor.  Checked here with power-of-two
    valued weights / inputs so that the bf16 rounding of the pre-summed taps is exact (the GPU test covers real values)."""
    import torch.nn.functional as F
    from faceposegenerator_b200.packing import pack_upsample_phase_weights
    g = torch.Generator().manual_seed(2)
    cin, cout, H, W = 8, 5, 6, 7
    w = torch.randint(-3, 4, (cout, cin, 3, 3), generator=g).float() / 4          # sums of <= 4 taps stay exact in bf16
    x = torch.randint(-8, 9, (2, cin, H, W), generator=g).float() / 8
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    wph = pack_upsample_phase_weights(w)
    out = torch.empty_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))                                                   # zero padding == TMA out-of-bounds fill
    for a in range(2):
        for c in range(2):
            k = wph[a][c].float().view(cout, 2, 2, cin).permute(0, 3, 1, 2)       # [Cout, 4*Cin] tap-major -> [Cout, Cin, 2, 2]
            # tap (u, v) reads low-resolution pixel (i + a - 1 + u, j + c - 1 + v)
            y = F.conv2d(xp[:, :, a:a + H + 1, c:c + W + 1], k)
            out[:, :, a::2, c::2] = y
    assert torch.equal(out, ref)
    # the one-launch form (IDB_EPI_PHASES4) takes the four matrices stacked on N, phase 2a + c major; the per-phase operands
    # are row-slice views of that tensor (no second copy of the weights)
    w_all, views = pack_upsample_phase_weights(w, stacked=True)
    assert tuple(w_all.shape) == (4 * cout, 4 * cin)
    for a in range(2):
        for c in range(2):
            assert torch.equal(views[a][c], wph[a][c]) and torch.equal(w_all[(2 * a + c) * cout:(2 * a + c + 1) * cout], wph[a][c])
            assert views[a][c].data_ptr() == w_all.data_ptr() + (2 * a + c) * cout * 4 * cin * w_all.element_size()


def test_iresnet_batchnorm_folding_is_exact_in_eval_mode():
    """Host-side weight preparation of the ArcFace backbone (iresnet.py): a conv followed by an eval-mode BatchNorm ==
    the conv with scaled weights + a bias (`_bn_affine` / `_fold_conv`), in the tap-major / channel-minor operand layout."""
    import torch.nn.functional as F
    from faceposegenerator_b200.iresnet import _bn_affine, _fold_conv
    g = torch.Generator().manual_seed(9)
    cin, cout = 6, 10
    w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
    sd = {"bn.weight": torch.rand(cout, generator=g, dtype=torch.float64) + 0.5, "bn.bias": torch.randn(cout, generator=g, dtype=torch.float64),
          "bn.running_mean": torch.randn(cout, generator=g, dtype=torch.float64), "bn.running_var": torch.rand(cout, generator=g, dtype=torch.float64) + 0.1}
    x = torch.randn(2, cin, 9, 9, generator=g, dtype=torch.float64)
    ref = F.batch_norm(F.conv2d(x, w, padding=1), sd["bn.running_mean"], sd["bn.running_var"], sd["bn.weight"], sd["bn.bias"],
                       training=False, eps=1e-5)
    scale, shift = _bn_affine(sd, "bn")
    wk, bias = _fold_conv(w, scale, shift)                                    # [Cout, 9 * Cin], tap-major
    w_back = wk.view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    got = F.conv2d(x, w_back, bias, padding=1)
    assert torch.allclose(got, ref, atol=1e-10)


def test_lora_file_round_trip(tmp_path):
    from faceposegenerator_b200 import weights as w
    lora = w.random_lora(seed=5)
    d = tmp_path / "checkpoint-31-6400"
    w.save_lora_weights(str(d), lora)
    back = w.load_lora_state(str(d))
    assert set(back) == set(lora) and len(back) == 128
    for k in lora:
        assert torch.equal(back[k][0], lora[k][0]) and torch.equal(back[k][1], lora[k][1]) and back[k][2] == 1.0
    # peft / legacy spellings and alpha scaling
    from safetensors.torch import save_file
    k0 = "down_blocks.0.attentions.0.transformer_blocks.0.attn1"
    t = {f"unet.{k0}.to_q.lora_A.weight": torch.ones(4, 320), f"unet.{k0}.to_q.lora_B.weight": torch.ones(320, 4),
         f"unet.{k0}.to_q.alpha": torch.tensor(8.0),
         f"unet.{k0}.processor.to_k_lora.down.weight": torch.ones(4, 320),
         f"unet.{k0}.processor.to_k_lora.up.weight": torch.ones(320, 4),
         f"unet.{k0}.to_out.0.lora.down.weight": torch.ones(4, 320), f"unet.{k0}.to_out.0.lora.up.weight": torch.ones(320, 4),
         "text_encoder.text_model.encoder.layers.0.self_attn.q_proj.lora.down.weight": torch.ones(4, 8)}
    save_file(t, str(tmp_path / "alt.safetensors"))
    alt = w.load_lora_state(str(tmp_path / "alt.safetensors"))
    assert set(alt) == {k0 + ".to_q", k0 + ".to_k", k0 + ".to_out.0"}
    assert alt[k0 + ".to_q"][2] == 2.0 and alt[k0 + ".to_k"][2] == 1.0
    with pytest.raises(FileNotFoundError):
        w.load_lora_state(str(tmp_path / "missing"))


def test_packing_layouts():
    from faceposegenerator_b200.packing import interleave_geglu, pack_conv_weight, pack_lora
    w = torch.arange(64 * 2 * 3 * 3, dtype=torch.float32).reshape(2, 64, 3, 3) / 1000
    p = pack_conv_weight(w)
    assert p.shape == (2, 576) and p.dtype == torch.bfloat16
    assert torch.equal(p[1, 5 * 64 + 7], w[1, 7, 1, 2].to(torch.bfloat16))   # tap (dy=1,dx=2) = 5
    wg = torch.arange(64.0)[:, None].repeat(1, 4)
    wi, bi = interleave_geglu(wg, torch.arange(64.0))
    assert wi[:, 0].tolist()[:48] == list(range(16)) + list(range(32, 48)) + list(range(16, 32))
    assert torch.equal(bi, wi[:, 0])
    ld, lu = pack_lora([(torch.ones(4, 64), torch.ones(320, 4), 2.0), None], seg_n=320, k=64)
    assert ld.shape == (32, 64) and lu.shape == (640, 64) and lu.dtype == torch.bfloat16
    assert float(ld[:4].float().sum()) == 256 and float(ld[4:].float().abs().sum()) == 0
    assert float(lu[:320, :4].float().sum()) == 2.0 * 320 * 4 and float(lu[320:].float().abs().sum()) == 0
    assert float(lu[:, 4:].float().abs().sum()) == 0
    assert pack_lora([None, None]) == (None, None)


def test_reference_script_imports_resolve():
    """Every import at /root/reference/inference_ID-Booth.py:1-15 resolves through compat/ (no
    GPU needed), and the pipeline object refuses to run without CUDA instead of falling back."""
    import subprocess
    code = ("from diffusers import StableDiffusionPipeline, DPMSolverMultistepScheduler, DDPMScheduler, "
            "AutoPipelineForText2Image\nfrom accelerate.utils import set_seed\nimport torch\n"
            "from torchvision.utils import save_image\nset_seed(0)\n"
            "p = StableDiffusionPipeline.from_pretrained('stabilityai/stable-diffusion-2-1-base', torch_dtype=torch.float16)\n"
            "p.scheduler = DDPMScheduler.from_pretrained('stabilityai/stable-diffusion-2-1-base', subfolder='scheduler')\n"
            "p.set_progress_bar_config(disable=True)\nprint('ok', type(p.scheduler).__name__)\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "ok DDPMScheduler" in r.stdout, r.stderr[-2000:]


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from faceposegenerator_b200 import StableDiffusionPipeline, ops
    p = StableDiffusionPipeline.from_pretrained("stabilityai/stable-diffusion-2-1-base")
    with pytest.raises(RuntimeError):
        p.to("cuda:0")
    with pytest.raises(RuntimeError):
        p(prompt="x")
    with pytest.raises(RuntimeError):
        ops.layernorm(torch.zeros(4, 8), torch.ones(8), torch.zeros(8))


def test_integration_stub_matches_the_binding():
    """The ctypes struct a maintainer copies out of INTEGRATION.md has exactly the fields of the shipped binding, in order
    (a shorter struct would make the library read past it: VERDICT r1 weak #12)."""
    import ctypes as C
    from faceposegenerator_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    body = text[text.index("class GemmConvArgs(C.Structure):"):text.index("_lib.idb_gemm_conv.argtypes")]
    fields = re.findall(r'\("(\w+)", C\.(c_\w+)\)', body)
    assert [(n, getattr(C, t)) for n, t in fields] == list(_lib.GemmConvArgs._fields_)
    assert "idb_sizeof_args(0) == C.sizeof(GemmConvArgs)" in text


def test_weight_policy_and_snapshot_resolution(tmp_path, monkeypatch):
    """Random-init stand-ins are an explicit opt-in; hub ids resolve through the local Hugging Face cache."""
    from faceposegenerator_b200 import pipeline as pl
    monkeypatch.delenv("IDB_ALLOW_RANDOM_WEIGHTS", raising=False)
    assert pl.random_weights_allowed() is False and pl.random_weights_allowed(True) is True
    monkeypatch.setenv("IDB_ALLOW_RANDOM_WEIGHTS", "1")
    assert pl.random_weights_allowed() is True and pl.random_weights_allowed(False) is False
    snap = tmp_path / "hub" / "models--org--name" / "snapshots" / "abc123"
    (snap / "unet").mkdir(parents=True)
    (tmp_path / "hub" / "models--org--name" / "refs").mkdir()
    (tmp_path / "hub" / "models--org--name" / "refs" / "main").write_text("abc123")
    monkeypatch.setenv("HF_HUB_CACHE", str(tmp_path / "hub"))
    assert pl.resolve_snapshot("org/name") == str(snap)
    assert pl.resolve_snapshot(str(snap)) == str(snap)
    assert pl.resolve_snapshot("org/other") is None
    # .bin checkpoints are read, not silently replaced
    import torch
    torch.save({"w": torch.ones(2)}, snap / "unet" / "diffusion_pytorch_model.bin")
    sd = pl._load_component_state(str(snap), "unet")
    assert sd is not None and torch.equal(sd["w"], torch.ones(2))
    assert pl._load_component_state(str(snap), "vae") is None


def test_c_abi_library_exports_every_declared_symbol():
    """include/idb.h <-> libidb_b200.so <-> the ctypes binding agree (no compute calls)."""
    import ctypes
    from faceposegenerator_b200 import _lib
    from faceposegenerator_b200.csrc import build
    lib_path = build.build()
    header = open(os.path.join(ROOT, "include", "idb.h")).read()
    declared = set(re.findall(r"^\s*(?:int|size_t|uint64_t)\s+(idb_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().idb_version() == 100
    # struct layouts: what gcc computes from include/idb.h == what the ctypes binding assumes
    import subprocess
    import tempfile
    structs = {"idb_gemm_conv_args": _lib.GemmConvArgs, "idb_attention_args": _lib.AttentionArgs,
               "idb_groupnorm_args": _lib.GroupNormArgs, "idb_time_embed_args": _lib.TimeEmbedArgs,
               "idb_attention_bwd_args": _lib.AttentionBwdArgs, "idb_groupnorm_bwd_args": _lib.GroupNormBwdArgs}
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "idb.h"\nint main(void){\n'
    for cname, cls in structs.items():
        prog += f'printf("{cname} %zu\\n", sizeof({cname}));\n'
        for fname, _ in cls._fields_:
            prog += f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));\n'
    enums = {"IDB_EPI_GEGLU": _lib.EPI_GEGLU, "IDB_EPI_F16": _lib.EPI_F16, "IDB_EPI_GELU": _lib.EPI_GELU, "IDB_EPI_PHASES4": _lib.EPI_PHASES4,
             "IDB_A_1X1": _lib.A_1X1, "IDB_A_3X3": _lib.A_3X3, "IDB_A_3X3_S2": _lib.A_3X3_S2, "IDB_A_3X3_S2_ASYM": _lib.A_3X3_S2_ASYM,
             "IDB_A_2X2": _lib.A_2X2}
    for ename in enums:
        prog += f'printf("{ename} %d\\n", (int){ename});\n'
    prog += "return 0;}\n"
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "l.c"), os.path.join(td, "l")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        got = dict(line.split() for line in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for ename, val in enums.items():
        assert int(got[ename]) == val, ename
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_oracle_is_not_imported_by_the_product():
    pkg = os.path.join(ROOT, "faceposegenerator_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py") and fn != "selfcheck.py":
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the reference path restated on host cores) must print exactly one JSON line with
    the contract's keys, also without a GPU."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
