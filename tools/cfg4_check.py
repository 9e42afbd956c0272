"""Config 4 (SD2.1 768 x 768: 96 x 96 latents): UNet forward against the fp32 oracle with per-block taps, VAE decode of a
96 x 96 latent, and the step time at B = 16.  NOT yet run on a GPU (written when the round's GPU budget was spent, see
DESIGN.md "Config 4 -- status"); run it first next round:
    gpurun --timeout 600 -- 'python tools/cfg4_check.py > gpurun_out/cfg4_check.log 2>&1; tail -20 gpurun_out/cfg4_check.log'
usage: python tools/cfg4_check.py [--no-oracle] [--bench-batch 16]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200.unet import UNet2DConditionModel  # noqa: E402
from faceposegenerator_b200.weights import random_lora, random_state_dict, unet_manifest, vae_decoder_manifest  # noqa: E402

FLOP_PER_ROW_96 = 2.1516e12     # UNet forward per batch row at 96 x 96


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--bench-batch", type=int, default=16)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    sd = random_state_dict(unet_manifest(), 0)
    lora = random_lora(seed=1)
    unet = UNet2DConditionModel(sd, device=dev)
    unet.set_lora(lora)
    g = torch.Generator().manual_seed(96)
    x = torch.randn(2, 4, 96, 96, generator=g)
    ctx = torch.randn(2, 77, 1024, generator=g)
    taps_g = {}
    out = unet.forward(x.to(dev), 496, ctx.to(dev), return_dict=False, taps=taps_g)[0]
    torch.cuda.synchronize()
    print(json.dumps({"unet_96x96_forward": "ok", "shape": list(out.shape), "finite": bool(torch.isfinite(out).all())}), flush=True)
    if not a.no_oracle:
        from oracle import sd21
        taps_o = {}
        t0 = time.time()
        with torch.no_grad():
            ref = sd21.unet_forward(sd, x, 496, ctx, lora, taps=taps_o)
        worst = max((rel(taps_g[k], taps_o[k]), k) for k in taps_o)
        print(json.dumps({"eps_rel_l2": rel(out, ref), "worst_tap": worst, "oracle_s": round(time.time() - t0, 1),
                          "tolerance": 1e-2}), flush=True)
    # VAE decode of a 96 x 96 latent (768 x 768 image)
    from faceposegenerator_b200.vae import AutoencoderKL
    from faceposegenerator_b200.weights import VAE_CONFIG
    vsd = random_state_dict(vae_decoder_manifest(VAE_CONFIG), 0)
    vae = AutoencoderKL(vsd, VAE_CONFIG, dev)
    z = torch.randn(1, 4, 96, 96, generator=g)
    img = vae.decode(z.to(dev), output_image=True)[0]
    torch.cuda.synchronize()
    line = {"vae_768_decode": "ok", "shape": list(img.shape), "finite": bool(torch.isfinite(img).all())}
    if not a.no_oracle:
        from oracle import sd21
        with torch.no_grad():
            ref_img = sd21.postprocess_np(sd21.vae_decode(vsd, z))
        d = img.float().cpu().numpy() - ref_img
        line["psnr_db"] = float(10 * torch.log10(torch.tensor(1.0 / max(float((d ** 2).mean()), 1e-20))))
    print(json.dumps(line), flush=True)
    # step time at the config's batch (CUDA graph, as in tools/unet_bench.py)
    B = a.bench_batch
    xb = torch.randn(B, 4, 96, 96, device=dev)
    cb = torch.randn(B, 77, 1024, device=dev)
    tb = torch.full((B,), 500.0, device=dev)
    context = unet.encode_context(cb)
    temb = unet.time_embedding(tb)
    for _ in range(2):
        unet.forward(xb, tb, context=context, temb=temb)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        unet.forward(xb, tb, context=context, temb=temb, return_dict=False)
    for _ in range(2):
        gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"B": B, "latent": "96x96", "graph_ms": round(ms, 3), "tflops": round(FLOP_PER_ROW_96 * B / ms / 1e9, 1)}))


if __name__ == "__main__":
    main()
