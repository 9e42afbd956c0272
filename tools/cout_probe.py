import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0")
x = torch.randn(4, 512, 512, 128, device=dev).bfloat16()
w = torch.randn(3, 3, 3, 128, device=dev) * 0.03
b = torch.randn(3, device=dev)
out = torch.empty(4, 512, 512, 3, device=dev)
fn = lambda: ops.conv3x3_small_cout(x, w, b, postprocess=True, out=out)
for _ in range(3): fn()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(5): fn()
g.replay(); torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): g.replay()
e.record(); torch.cuda.synchronize()
print(json.dumps({"conv_small_cout_512x512x128_B4_us": round(a.elapsed_time(e) / 25 * 1e3, 1)}))
