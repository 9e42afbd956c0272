"""ArcFace IResNet-100 + the x0 -> ArcFace glue on the CUDA path against the PINNED oracle
(oracle/iresnet.py, itself checked against outputs of the reference module: tests/golden/iresnet100_golden.pt)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.fixture(scope="module")
def ops(cuda_dev):
    from faceposegenerator_b200 import ops
    return ops


@pytest.fixture(scope="module")
def golden_sd():
    from faceposegenerator_b200.weights import iresnet_manifest
    from oracle.iresnet import keyed_state_dict
    shapes = dict(iresnet_manifest("r100"))
    return keyed_state_dict({k: torch.zeros(s, dtype=torch.long if k.endswith("tracked") else torch.float32)
                             for k, s in shapes.items()}, seed=0)


@pytest.fixture(scope="module")
def model(golden_sd, cuda_dev):
    from faceposegenerator_b200.iresnet import IResNet
    return IResNet(golden_sd, "r100", device=cuda_dev)


def test_iresnet100_vs_reference_golden(model, golden_sd, cuda_dev):
    """Embedding of the reference module's own output fixture (same weights, same seeded input).
    bf16 operands / fp32 accumulation through 100 conv layers: rel-L2 <= 1e-2 (north_star tolerance for bf16)."""
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "iresnet100_golden.pt"))
    x = torch.randn(2, 3, 112, 112, generator=torch.Generator().manual_seed(0))
    y = model(x.to(cuda_dev)).cpu()
    assert y.shape == (2, 512)
    r = rel(y, gold["embedding"])
    cos = F.cosine_similarity(y, gold["embedding"], dim=-1)
    print(f"iresnet100 embedding rel-L2 {r:.3e}, cosine {cos.tolist()}")
    assert r < 1e-2, r
    assert float(cos.min()) > 0.9999


def test_iresnet100_batch_and_oracle(model, golden_sd, cuda_dev):
    """Config-5 batch (4 images) against the fp32 oracle on fresh inputs; rows are independent."""
    from oracle.iresnet import iresnet_forward
    x = torch.randn(4, 3, 112, 112, generator=torch.Generator().manual_seed(7)).clamp(-1, 1)
    with torch.no_grad():
        ref = iresnet_forward(golden_sd, x)
    y = model(x.to(cuda_dev)).cpu()
    assert rel(y, ref) < 1e-2, rel(y, ref)
    y1 = model(x[2:3].to(cuda_dev)).cpu()
    assert rel(y1, y[2:3]) < 2e-3


def test_crop_resize_norm_vs_torch(ops, cuda_dev):
    """train_ID-Booth.py:1088-1092 + :445-455: crop the bbox, resize(112, antialias=None), (x/255 - 0.5)/0.5."""
    g = torch.Generator().manual_seed(3)
    img = torch.rand(3, 256, 256, 3, generator=g)
    boxes = torch.tensor([[48, 40, 208, 216], [-5, 10, 300, 250], [96, 96, 160, 161]], dtype=torch.int32)
    out = ops.crop_resize_norm(img.to(cuda_dev), boxes.to(cuda_dev), size=112, c_pad=64).float().cpu()
    assert out.shape == (3, 112, 112, 64) and float(out[..., 3:].abs().max()) == 0.0
    for i in range(3):
        x0, y0, x1, y1 = [int(v) for v in boxes[i]]
        im = (img[i] * 255)
        crop = im[max(0, y0):min(y1, 256), max(0, x0):min(x1, 256)]              # HWC, like the reference
        t = crop.permute(2, 0, 1)[None]
        t = F.interpolate(t, size=(112, 112), mode="bilinear", align_corners=False, antialias=False)[0]
        ref = ((t / 255) - 0.5) / 0.5
        got = out[i, :, :, :3].permute(2, 0, 1)
        assert float((got - ref).abs().max()) < 1e-2      # bf16 output rounding (values in [-1, 1])
        assert rel(got, ref) < 4e-3


def test_crop_resize_norm_vs_reference_golden(ops, cuda_dev):
    """`idb_crop_resize_norm` against outputs of the reference's own glue functions (train_ID-Booth.py:433-455,1090;
    tests/golden/arcface_glue_golden.pt): decoded image -> [0, 1] NHWC (the VAE's fused post-process) -> crop + bilinear
    112 x 112 + normalise in one kernel."""
    import os
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "arcface_glue_golden.pt"))
    for k, (seed, size, bbox) in enumerate(gold["cases"]):
        dec = torch.rand(1, 3, size, size, generator=torch.Generator().manual_seed(seed)) * 2.4 - 1.2
        img01 = (dec / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).contiguous()
        out = ops.crop_resize_norm(img01.to(cuda_dev), torch.tensor([bbox], dtype=torch.int32, device=cuda_dev),
                                   size=112, c_pad=64).float().cpu()
        got = out[0, :, :, :3].permute(2, 0, 1)
        want = gold["arcface_inputs"][k]
        if want.shape[-1] != 112:
            got = got[:, ::4, ::4]
        assert float((got - want).abs().max()) < 1e-2 and rel(got, want) < 4e-3   # bf16 output rounding on [-1, 1]


def test_channel_affine(ops, cuda_dev):
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(3, 14, 14, 256, device=cuda_dev, generator=g)
    sc = torch.randn(256, device=cuda_dev, generator=g)
    sh = torch.randn(256, device=cuda_dev, generator=g)
    y = ops.channel_affine(x, sc, sh).float()
    assert rel(y, x * sc + sh) < 4e-3
    y2 = ops.channel_affine(x, stride=2).float()
    assert y2.shape == (3, 7, 7, 256) and rel(y2, x[:, ::2, ::2]) < 4e-3


def test_identity_loss_end_to_end(model, golden_sd, cuda_dev):
    """x0-image -> ArcFace embedding -> 1 - cos (train_ID-Booth.py:1093-1098) with a fixed bbox (SURVEY 8d cfg 5)."""
    from faceposegenerator_b200.iresnet import arcface_embedding_from_images, identity_loss
    from oracle.iresnet import iresnet_forward
    g = torch.Generator().manual_seed(11)
    img = torch.rand(2, 512, 512, 3, generator=g)
    bbox = torch.tensor([[96, 96, 416, 416]] * 2, dtype=torch.int32)
    emb = arcface_embedding_from_images(model, img.to(cuda_dev), bbox.to(cuda_dev)).cpu()
    crop = (img * 255)[:, 96:416, 96:416].permute(0, 3, 1, 2)
    xin = ((F.interpolate(crop, size=(112, 112), mode="bilinear", align_corners=False) / 255) - 0.5) / 0.5
    with torch.no_grad():
        ref = iresnet_forward(golden_sd, xin)
    assert rel(emb, ref) < 1.5e-2
    gt = ref[[1, 0]]
    assert torch.allclose(identity_loss(emb, gt), identity_loss(ref, gt), atol=5e-3)


def test_config5_training_forward_identity(model, golden_sd, cuda_dev):
    """Config 5 at its stated batch (B = 4, no CFG, `train_ID-Booth.py:1040-1046,1109-1133`): UNet(noisy, t, ctx) ->
    pred_original_sample -> VAE decode -> crop -> 112x112 -> IResNet, against the same chain in the fp32 oracles."""
    from faceposegenerator_b200 import DDPMScheduler
    from faceposegenerator_b200.iresnet import training_forward_identity
    from faceposegenerator_b200.unet import UNet2DConditionModel
    from faceposegenerator_b200.vae import AutoencoderKL
    from faceposegenerator_b200.weights import random_lora, random_state_dict, unet_manifest, vae_decoder_manifest
    from oracle import sd21
    from oracle.iresnet import iresnet_forward
    usd, vsd = random_state_dict(unet_manifest(), 0), random_state_dict(vae_decoder_manifest(), 0)
    lora = random_lora(seed=1)
    unet = UNet2DConditionModel(usd, device=cuda_dev)
    unet.set_lora(lora)
    vae = AutoencoderKL(vsd, device=cuda_dev)
    sched = DDPMScheduler.from_pretrained("stabilityai/stable-diffusion-2-1-base", subfolder="scheduler")
    g = torch.Generator().manual_seed(0)
    nb = 4
    x0_true = torch.randn(nb, 4, 64, 64, generator=g)
    noise = torch.randn(nb, 4, 64, 64, generator=g)
    ctx = torch.randn(nb, 77, 1024, generator=g)
    ts = [437, 112, 871, 3]
    ref_s = sd21.DDPMSchedulerRef()
    noisy = ref_s.add_noise(x0_true, noise, torch.tensor(ts))
    bbox = torch.tensor([[96, 96, 416, 416]] * nb, dtype=torch.int32)
    eps, x0, emb = training_forward_identity(unet, vae, sched, model, noisy.to(cuda_dev), ts, ctx.to(cuda_dev), bbox.to(cuda_dev))
    with torch.no_grad():
        eps_r = torch.cat([sd21.unet_forward(usd, noisy[i:i + 1], ts[i], ctx[i:i + 1], lora) for i in range(nb)])
        x0_r = torch.cat([ref_s.step(eps_r[i:i + 1], ts[i], noisy[i:i + 1], torch.zeros(1, 4, 64, 64))[1] for i in range(nb)])
        img_r = (sd21.vae_decode(vsd, x0_r / 0.18215) * 0.5 + 0.5).clamp(0, 1)                 # NCHW
        crop = (img_r * 255)[:, :, 96:416, 96:416]
        xin = ((F.interpolate(crop, size=(112, 112), mode="bilinear", align_corners=False) / 255) - 0.5) / 0.5
        emb_r = iresnet_forward(golden_sd, xin)
    print(f"cfg5: eps {rel(eps.cpu(), eps_r):.3e} x0 {rel(x0.cpu(), x0_r):.3e} emb {rel(emb.cpu(), emb_r):.3e} "
          f"cos {F.cosine_similarity(emb.cpu(), emb_r, dim=-1).tolist()}")
    assert rel(eps.cpu(), eps_r) < 1e-2
    assert rel(x0.cpu(), x0_r) < 1e-2
    assert float(F.cosine_similarity(emb.cpu(), emb_r, dim=-1).min()) > 0.995


def test_extract_embeds_on_cuda(model, golden_sd, cuda_dev, tmp_path, monkeypatch):
    """`extract_embeds.run` (the reference's extract_ArcFace_embeds.py) with the CUDA backbone: per-folder embeddings of
    every detected face against the pinned CPU oracle on the same preprocessed crops; host logic is pinned on the CPU
    side (tests/test_extract_embeds_cpu.py)."""
    import json
    import numpy as np
    from PIL import Image
    from faceposegenerator_b200.extract_embeds import crop_to_bbox, prepare_for_arcface, run
    from oracle.iresnet import iresnet_forward
    monkeypatch.chdir(tmp_path)
    rng = np.random.RandomState(0)
    boxes = {}
    for folder, sizes in (("p1", (160, 160)), ("p2", (128,))):   # the images of a folder are stacked (`:46`): one size per folder
        os.makedirs(os.path.join("FACE_DATASET", "images", folder))
        for k, size in enumerate(sizes):
            arr = np.kron(rng.randint(0, 256, size=(size // 8, size // 8, 3)).astype(np.uint8), np.ones((8, 8, 1), np.uint8))
            Image.fromarray(arr).save(os.path.join("FACE_DATASET", "images", folder, f"{k}.png"))
            boxes[os.path.join("images", folder, f"{k}.png")] = [12 + 9 * k, 9, size - 17, size - 6 - 11 * k]
    with open("boxes.json", "w") as f:
        json.dump(boxes, f)
    without = run("FACE_DATASET", device="cuda:0", model=model, bbox_file="boxes.json", embed="all",
                  listdir=lambda p: sorted(os.listdir(p)))
    assert without == {"files_without_faces": []}
    for folder, n in (("p1", 2), ("p2", 1)):
        emb = torch.load(os.path.join("FACE_DATASET", "ArcFace_embeds", folder, folder + ".pt")).float().cpu()
        assert emb.shape == (n, 512)
        crops = []
        for k in range(n):
            rel_path = os.path.join("images", folder, f"{k}.png")
            img = torch.from_numpy(np.array(Image.open(os.path.join("FACE_DATASET", rel_path))))
            crops.append(prepare_for_arcface(crop_to_bbox(img, boxes[rel_path])))
        with torch.no_grad():
            ref = iresnet_forward(golden_sd, torch.cat(crops, 0))
        assert rel(emb, ref) < 1.5e-2, rel(emb, ref)
