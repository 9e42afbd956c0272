"""Per-phase clock() breakdown of the GEMM epilogue warps (IDB_GEMM_DEBUG=0x400 instrumentation)."""
import math, os, sys
os.environ["IDB_GEMM_DEBUG"] = str(0x400)
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16; f32 = torch.float32
NAMES = ["prologue", "wait_acc", "tmem_ld", "math", "res/wait_read", "staging", "fence", "store"]
def run(name, M, K, N, *, f32out=False, b16out=False, res_=False, stats=False, lora=0, geglu=False, bias=True):
    n_out = N // 2 if geglu else N
    x = torch.randn(M, K, device=dev).to(bf16)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf16)
    ws = torch.zeros(1 << 20, dtype=f32, device=dev)
    kw = dict(workspace=ws, k_splits=1)
    if bias: kw["bias"] = torch.randn(N, device=dev)
    if res_: kw["residual"] = torch.randn(M, n_out, device=dev)
    if f32out: kw["out_f32"] = torch.empty(M, n_out, dtype=f32, device=dev)
    if b16out: kw["out_bf16"] = torch.empty(M, n_out, dtype=bf16, device=dev)
    if stats: kw["stats"] = torch.empty((M + 31) // 32, n_out, 2, dtype=f32, device=dev)
    if lora:
        kw["lora_down"] = torch.randn(16 * lora, K, device=dev).to(bf16)
        up = torch.zeros(N, 64, device=dev); up[:, :4] = torch.randn(N, 4, device=dev) * 0.05
        kw["lora_up"] = up.to(bf16)
        kw["lora_seg_n"] = N // lora
    if geglu: kw["geglu"] = True
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        big.zero_()   # flush L2
        ws.zero_()
        ops.gemm_conv(x, w, **kw)
    torch.cuda.synchronize()
    raw = ws.view(torch.int32).cpu()
    nblk = 148
    epi = raw[: nblk * 16 * 8].view(nblk, 16, 8).float()
    mma = raw[nblk * 16 * 8: nblk * 16 * 8 + nblk * 4].view(nblk, 4).float()
    print(f"== {name}: M{M} K{K} N{N}")
    tot = epi.sum(-1)
    print("  epilogue warp total clk: mean %.0f max %.0f" % (tot.mean(), tot.max()))
    for k, nm in enumerate(NAMES):
        print("   %-14s mean %8.0f  (%.1f%%)" % (nm, epi[:, :, k].mean(), 100 * epi[:, :, k].mean() / tot.mean()))
    lead = mma[mma[:, 3] > 0]
    print("  MMA warp (leaders): tiles %.1f total %.0f wait_empty %.0f wait_full %.0f" % (lead[:, 3].mean(), lead[:, 2].mean(), lead[:, 0].mean(), lead[:, 1].mean()))
sel = sys.argv[1:] or ["proj_in", "proj_out", "qkv_plain", "ff1"]
if "proj_in" in sel: run("proj_in", 32768, 320, 320, f32out=True)
if "proj_out" in sel: run("proj_out", 32768, 320, 320, f32out=True, res_=True, stats=True)
if "qkv_plain" in sel: run("qkv_plain", 32768, 320, 960, b16out=True, bias=False)
if "qkv_lora" in sel: run("qkv_lora", 32768, 320, 960, b16out=True, bias=False, lora=3)
if "ff1" in sel: run("ff1", 32768, 320, 2560, b16out=True, geglu=True)
if "ff2" in sel: run("ff2", 32768, 1280, 320, b16out=True, res_=True)
