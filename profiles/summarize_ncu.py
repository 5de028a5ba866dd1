#!/usr/bin/env python
"""Turn an Nsight Compute report (+ optional launch-list CSV) into the short text summaries kept in
profiles/.  Runs in the build container (no GPU needed):

    python profiles/summarize_ncu.py gpurun_out/prof_r1c.ncu-rep [gpurun_out/launches_r1c.csv] > profiles/r1c_summary.txt
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

RAW_KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except Exception:
        return 0.0


def raw_summary(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"== {r[kn][:90]}")
        for k in RAW_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:72s} {r[i]:>18s} {units[i]}")
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        def to_bytes(v, u):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            return num(v) * scale
        print(f"   traffic (dram read+write bytes per launch)                               "
              f"{to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]):18.0f} byte")


def stall_summary(rep, kernel_regex):
    rows = ncu_csv(rep, "source", ["--kernel-name", f"regex:{kernel_regex}"])
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hi:
        return
    end = hi[1] if len(hi) > 1 else len(rows)          # several kernels may match: keep the first one's rows
    hdr = rows[hi[0]]
    data = [r for r in rows[hi[0] + 1:end] if len(r) == len(hdr)]
    si, ii = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(num(r[si]) for r in data) or 1.0
    agg = sorted(((hdr[c][6:], sum(num(r[c]) for r in data)) for c in cols), key=lambda kv: -kv[1])
    print(f"== stall samples, {kernel_regex}: total {tot:.0f}; instructions executed "
          f"{sum(num(r[ii]) for r in data) / 1e6:.0f} M (SASS lines {len(data)})")
    print("   by reason: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in agg[:8]))
    print("   hottest SASS lines:")
    for r in sorted(data, key=lambda r: -num(r[si]))[:12]:
        print(f"     {100 * num(r[si]) / tot:5.1f}%  exec {num(r[ii]) / 1e6:8.2f}M  {r[1].strip()[:80]}")


def launch_list(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in data:
        if len(r) > mv:
            agg[r[kn][:70]].append(num(r[mv]))
    total = sum(sum(v) for v in agg.values()) or 1.0
    print(f"== launch list {path}: {sum(len(v) for v in agg.values())} launches, {total / 1e6:.2f} ms total "
          "(cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:14]:
        print(f"   {100 * sum(v) / total:5.1f}%  n={len(v):4d}  avg {sum(v) / len(v) / 1e3:9.1f} us  {k}")


if __name__ == "__main__":
    rep = sys.argv[1]
    raw_summary(rep)
    for rx in ("dense_pair", "dense_umma", "bm25_select"):
        stall_summary(rep, rx)
    if len(sys.argv) > 2:
        launch_list(sys.argv[2])
