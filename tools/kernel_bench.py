"""Per-kernel micro-benchmarks on one B200 (CUDA events, L2 flushed between iterations).
Prints one JSON line per case to stdout: achieved TFLOP/s (tensor-bound) or GB/s (HBM-bound)."""
import json
import math
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faceposegenerator_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf16, f32 = torch.bfloat16, torch.float32
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
WS = torch.empty((96 << 20) // 4, dtype=f32, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, flops=None, bytes_=None, **kw):
    d = {"case": name, "ms": round(ms, 4)}
    if flops:
        d["tflops"] = round(flops / ms / 1e9, 1)
    if bytes_:
        d["gbs"] = round(bytes_ / ms / 1e6, 1)
    d.update(kw)
    print(json.dumps(d), flush=True)


def bench_linear(M, K, N, geglu=False, lora=False, tag="linear"):
    x = torch.randn(M, K, device=dev).to(bf16)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf16)
    bias = torch.randn(N, device=dev)
    kw = {}
    if lora:
        kw = dict(lora_down=torch.randn(16 * (N // 320 if N % 320 == 0 and N > 1280 else 1), K, device=dev).to(bf16),
                  lora_up=torch.nn.functional.pad(torch.randn(N, 4, device=dev) * 0.05, (0, 60)).to(bf16),
                  lora_seg_n=N if N <= 1280 else N // 3)
    out = torch.empty(M, N // 2 if geglu else N, dtype=bf16, device=dev)
    ms = timeit(lambda: ops.gemm_conv(x, w, bias=bias, geglu=geglu, out_bf16=out, k_splits=0, workspace=WS, **kw))
    report(f"{tag} M{M} K{K} N{N}" + (" geglu" if geglu else "") + (" lora" if lora else ""), ms, flops=2.0 * M * K * N)


def bench_conv(B, H, C, N, stride=1, c1=0):
    x = torch.randn(B, H, H, C, device=dev).to(bf16)
    w = (torch.randn(N, 9 * C + c1, device=dev) / math.sqrt(9 * C)).to(bf16)
    a1 = torch.randn(B, H, H, c1, device=dev).to(bf16) if c1 else None
    bias = torch.randn(N, device=dev)
    Ho = H // stride
    out = torch.empty(B * Ho * Ho, N, dtype=f32, device=dev)
    mode = ops.A_3X3_S2 if stride == 2 else ops.A_3X3
    ms = timeit(lambda: ops.gemm_conv(x, w, mode=mode, a1=a1, bias=bias, out_f32=out, k_splits=0, workspace=WS))
    report(f"conv3x3 B{B} {H}x{H} {C}->{N} s{stride} sc{c1}", ms, flops=2.0 * B * Ho * Ho * N * (9 * C + c1))


def bench_attn(B, heads, T, S):
    C = heads * 64
    qkv = torch.randn(B * T, 3 * C, device=dev).to(bf16)
    kv = torch.randn(B * S, 2 * C, device=dev).to(bf16)
    out = torch.empty(B * T, C, dtype=bf16, device=dev)
    if T == S:
        fn = lambda: ops.attention(qkv, qkv, qkv, out, batch=B, heads=heads, t_q=T, t_kv=S, scale=0.125, col0_k=C, col0_v=2 * C)
    else:
        fn = lambda: ops.attention(qkv, kv, kv, out, batch=B, heads=heads, t_q=T, t_kv=S, scale=0.125, col0_v=C)
    ms = timeit(fn)
    report(f"attention B{B} h{heads} T{T} S{S}", ms, flops=4.0 * B * heads * T * S * 64)


def bench_gn(B, HW, C0, C1=0):
    x0 = torch.randn(B, HW, C0, device=dev)
    x1 = torch.randn(B, HW, C1, device=dev) if C1 else None
    C = C0 + C1
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    out = torch.empty(B, HW, C, dtype=bf16, device=dev)
    ws = ops.groupnorm_workspace(B, 32, dev)
    ms = timeit(lambda: ops.groupnorm(x0, g, b, groups=32, eps=1e-5, silu=True, x1=x1, out_norm=out, partials=ws))
    report(f"groupnorm B{B} HW{HW} C{C0}+{C1}", ms, bytes_=B * HW * C * (4 + 2))


def bench_ln(rows, C):
    x = torch.randn(rows, C, device=dev)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    out = torch.empty(rows, C, dtype=bf16, device=dev)
    ms = timeit(lambda: ops.layernorm(x, g, b, out))
    report(f"layernorm rows{rows} C{C}", ms, bytes_=rows * C * 6)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    for (H, C, N, s, c1) in [(64, 320, 320, 1, 0), (64, 640, 320, 1, 0), (64, 320, 320, 1, 960), (64, 640, 640, 1, 0),
                             (32, 640, 640, 1, 0), (32, 1280, 640, 1, 0), (32, 640, 640, 1, 1920), (32, 1280, 1280, 1, 0),
                             (16, 1280, 1280, 1, 0), (16, 2560, 1280, 1, 0), (8, 1280, 1280, 1, 0), (8, 2560, 1280, 1, 0),
                             (64, 320, 320, 2, 0), (32, 640, 640, 2, 0), (16, 1280, 1280, 2, 0)]:
        bench_conv(B, H, C, N, s, c1)
    for (hw, C) in [(4096, 320), (1024, 640), (256, 1280), (64, 1280)]:
        M = B * hw
        bench_linear(M, C, C, tag="proj")
        bench_linear(M, C, C, lora=True, tag="q/o")
        bench_linear(M, C, 3 * C, lora=True, tag="qkv")
        bench_linear(M, C, 8 * C, geglu=True, tag="ff1")
        bench_linear(M, 4 * C, C, tag="ff2")
    for (h, T) in [(5, 4096), (10, 1024), (20, 256), (20, 64)]:
        bench_attn(B, h, T, T)
        bench_attn(B, h, T, 77)
    bench_gn(B, 4096, 320); bench_gn(B, 4096, 640, 320); bench_gn(B, 1024, 640); bench_gn(B, 256, 1280, 1280); bench_gn(B, 64, 1280)
    bench_ln(B * 4096, 320); bench_ln(B * 1024, 640); bench_ln(B * 256, 1280)


if __name__ == "__main__":
    main()
