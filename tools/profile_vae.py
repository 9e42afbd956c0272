"""One eager VAE decode (B images, 64x64 latents -> 512x512) between cudaProfilerStart / Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --csv`; also prints the CUDA-event time of a graph replay."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200.vae import AutoencoderKL  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
vae = AutoencoderKL.from_random(0, device=dev)
z = torch.randn(B, 4, 64, 64, device=dev)
for _ in range(2):
    vae.decode(z, output_image=True)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    img = vae.decode(z, output_image=True)[0]
g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    g.replay()
b.record()
torch.cuda.synchronize()
print(json.dumps({"B": B, "vae_decode_graph_ms": round(a.elapsed_time(b) / 5, 3), "tflops": round(2.5145 * B / (a.elapsed_time(b) / 5) * 1e3, 1)}))
torch.cuda.profiler.start()
vae.decode(z, output_image=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
