"""Config 5 (BASELINE.json configs[4]): training-mode UNet forward, batch 4 (no CFG), + ArcFace IResNet-100 embedding of
the predicted x0 (VAE decode -> crop -> 112x112 -> backbone).  Random-init weights, synthetic inputs, fixed bbox."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import DDPMScheduler
from faceposegenerator_b200.iresnet import IResNet, arcface_embedding_from_images, identity_step_tables, training_forward_identity
from faceposegenerator_b200.unet import UNet2DConditionModel
from faceposegenerator_b200.vae import AutoencoderKL
from faceposegenerator_b200.weights import random_iresnet_state_dict, random_lora

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
unet = UNet2DConditionModel.from_random(0, device=dev); unet.set_lora(random_lora(seed=0))
vae = AutoencoderKL.from_random(0, device=dev) if hasattr(AutoencoderKL, "from_random") else None
arc = IResNet(random_iresnet_state_dict("r100", 0), "r100", device=dev)
sched = DDPMScheduler.from_pretrained("stabilityai/stable-diffusion-2-1-base", subfolder="scheduler")
g = torch.Generator(device="cuda").manual_seed(0)
noisy = torch.randn(B, 4, 64, 64, device=dev, generator=g)
ctx = torch.randn(B, 77, 1024, device=dev, generator=g)
ts = [int(t) for t in torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(0))]
bbox = torch.tensor([[96, 96, 416, 416]] * B, dtype=torch.int32, device=dev)
context = unet.encode_context(ctx)
def ev(): return torch.cuda.Event(enable_timing=True)
def timed(fn, n=5):
    for _ in range(2): fn()
    a, b = ev(), ev(); torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
res = {"config": f"cfg5: UNet fwd B={B} (no CFG) + x0 -> VAE decode -> crop/resize -> IResNet-100", "timesteps": ts}
res["chain_ms"] = round(timed(lambda: training_forward_identity(unet, vae, sched, arc, noisy, ts, ctx, bbox, context=context)), 3)
tt = torch.tensor(ts, device=dev, dtype=torch.float32)
res["unet_fwd_ms"] = round(timed(lambda: unet.forward(noisy, tt, context=context)), 3)
img = torch.rand(B, 512, 512, 3, device=dev)
res["arcface_ms"] = round(timed(lambda: arcface_embedding_from_images(arc, img, bbox)), 3)
lat = torch.randn(B, 4, 64, 64, device=dev)
res["vae_decode_ms"] = round(timed(lambda: vae.decode(lat, output_image=True)), 3)
res["iresnet_gflop_per_image"] = 24.2
res["iresnet_tflops"] = round(24.2e9 * B / (res["arcface_ms"] * 1e-3) / 1e12, 1)
# the same chain / backbone as ONE CUDA-graph launch (static shapes; the eager numbers above are bound by the host launch rate)
def graphed(fn):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    return gr.replay
try:
    tables = identity_step_tables(sched, ts, dev)
    res["chain_graph_ms"] = round(timed(graphed(lambda: training_forward_identity(unet, vae, sched, arc, noisy, None, ctx, bbox, context=context, tables=tables))), 3)
    res["arcface_graph_ms"] = round(timed(graphed(lambda: arcface_embedding_from_images(arc, img, bbox))), 3)
    res["iresnet_graph_tflops"] = round(24.2e9 * B / (res["arcface_graph_ms"] * 1e-3) / 1e12, 1)
except Exception as e:   # (a host synchronisation inside the chain would make it uncapturable)
    res["graph_error"] = f"{type(e).__name__}: {e}"[:200]
print(json.dumps(res))
