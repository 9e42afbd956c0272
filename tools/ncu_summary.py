"""Condense an `ncu --set full` report into one line per launch (and a per-kernel aggregate):
duration, tensor-pipe utilisation, DRAM bytes and throughput, registers.
usage: python tools/ncu_summary.py report.ncu-rep out.csv"""
import collections
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = {
    "dur_us": "gpu__time_duration.sum",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_pct_elapsed": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
    "dram_rate": "dram__bytes.sum.per_second",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
}
scale = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
         "byte/s": 1.0, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}


def val(r, key):
    name = want[key]
    if name not in ix:
        return float("nan")
    s = r[ix[name]].replace(",", "")
    try:
        v = float(s)
    except ValueError:
        return float("nan")
    return v * scale.get(units[ix[name]], 1.0)


agg = collections.OrderedDict()
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "dur_us", "tensor_pct_active", "tensor_pct_elapsed", "dram_read_MB", "dram_write_MB",
                "dram_GBps", "dram_pct", "sm_pct", "regs"])
    for r in data:
        k = r[ix["Kernel Name"]][:60]
        d = val(r, "dur_us")
        rd, wr = val(r, "dram_read"), val(r, "dram_write")
        if rd != rd:   # sections without the split counters: total bytes = rate x duration
            rd, wr = val(r, "dram_rate") * d * 1e-6, 0.0
        w.writerow([r[ix["ID"]], k, r[ix["Grid Size"]], f"{d:.1f}", f"{val(r, 'tensor_pct'):.1f}", f"{val(r, 'tensor_pct_elapsed'):.1f}",
                    f"{rd / 1e6:.2f}", f"{wr / 1e6:.2f}", f"{(rd + wr) / d / 1e3:.0f}" if d > 0 else "", f"{val(r, 'dram_pct'):.1f}",
                    f"{val(r, 'sm_pct'):.1f}", f"{val(r, 'regs'):.0f}"])
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d
        a[2] += val(r, "tensor_pct_elapsed") * d
        a[3] += rd + wr
        a[4] += val(r, "dram_pct") * d
    w.writerow([])
    w.writerow(["# per-kernel aggregate: kernel", "launches", "total_us", "time-weighted tensor_pct_elapsed", "dram_MB", "mean dram GB/s",
                "time-weighted dram_pct"])
    for k, (n, d, t, b, dp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow(["#", k, n, f"{d:.1f}", f"{t / d:.1f}" if d else "", f"{b / 1e6:.1f}", f"{b / d / 1e3:.0f}" if d else "", f"{dp / d:.1f}" if d else ""])
print(open(out).read()[-2500:])
