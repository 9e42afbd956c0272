"""Join an ncu launch list of ONE UNet forward (tools/profile_step.py) with its shape trace:
per-GEMM / per-attention achieved TFLOP/s.  usage: python tools/join_trace.py launches.csv step_trace.json"""
import collections
import csv
import json
import sys

with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
launches = []
for row in csv.DictReader(lines):
    if row["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    launches.append((row["Kernel Name"], row["Grid Size"], v))
# one forward = the launches from one latent_operand_kernel (conv_in operand) to the next; take the last complete one
starts = [i for i, l in enumerate(launches) if "latent_operand_kernel" in l[0]]
if len(starts) >= 2:
    launches = launches[starts[-2]:starts[-1]]
print(f"forward segment: {len(launches)} launches, {sum(l[2] for l in launches):.1f} us")
trace = json.load(open(sys.argv[2]))
gem = [l for l in launches if "gemm_tc_kernel" in l[0]]
att = [l for l in launches if "attention" in l[0]]
tg = [d for n, d in trace if n == "idb_gemm_conv"]
ta = [d for n, d in trace if n == "idb_attention"]
print(f"gemm launches {len(gem)} vs trace {len(tg)}; attention {len(att)} vs {len(ta)}")
groups = collections.OrderedDict()
for (name, grid, us), d in zip(gem, tg):
    fl = 2.0 * d["M"] * d["N"] * d["K"]
    key = (d["M"], d["N"], d["K"], d["mode"], d["lora"], d["geglu"], d["f32"], d["stats"], name.split("<")[1].split(">")[0])
    g = groups.setdefault(key, [0, 0.0, 0.0])
    g[0] += 1
    g[1] += us
    g[2] += fl
print(f"{'M':>6} {'N':>5} {'K':>6} md lora geglu f32 st {'kernel':>12} {'n':>3} {'us/launch':>9} {'tot us':>8} {'TFLOP/s':>8}")
tot_us = tot_fl = 0
for k, (n, us, fl) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:6d} {k[1]:5d} {k[2]:6d} {k[3]:2d} {int(k[4]):4d} {int(k[5]):5d} {int(k[6]):3d} {int(k[7]):2d} {k[8]:>12} {n:3d} {us / n:9.1f} {us:8.1f} {fl / us / 1e6:8.1f}")
    tot_us += us
    tot_fl += fl
print(f"GEMM total {tot_us:.1f} us, {tot_fl / 1e12:.3f} TFLOP, {tot_fl / tot_us / 1e6:.1f} TFLOP/s")
ag = collections.OrderedDict()
for (name, grid, us), d in zip(att, ta):
    fl = 4.0 * d["B"] * d["heads"] * d["Tq"] * d["Tkv"] * 64
    g = ag.setdefault((d["B"], d["heads"], d["Tq"], d["Tkv"], name[:24]), [0, 0.0, 0.0])
    g[0] += 1
    g[1] += us
    g[2] += fl
for k, (n, us, fl) in ag.items():
    print(f"attn B{k[0]} h{k[1]} Tq{k[2]} Tkv{k[3]} {k[4]}: n={n} {us / n:.1f} us/launch, total {us:.1f} us, {fl / us / 1e6:.1f} TFLOP/s")
