#!/usr/bin/env python
"""Per-source-line view of an Nsight Compute capture: joins the SASS page of a kernel (sampled stalls,
instructions executed) with the line table nvdisasm prints for the SAME build of libdrs_b200.so.
Instructions are matched by their index inside the kernel, so the .so must be the one profiled.

    python profiles/line_profile.py gpurun_out/prof_r1g.ncu-rep bm25_select bm25 [top_n] [mangled-regex]
                                    report                       kernel-regex cubin-stem
(mangled-regex picks the template instance in the cubin, e.g. bm25_select_kernelILb0 for <false>)
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "domain_rag_system_b200", "csrc", "libdrs_b200.so")


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except Exception:
        return 0.0


def sass_lines(cubin_stem, kernel_regex):
    """[(line_no, inline_chain, sass_text)] in instruction order for the first kernel matching the regex."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_stem + ".") and f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    out, on, line = [], False, 0
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            on = re.search(kernel_regex, ln) is not None
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]*)", line (\d+)', ln)
        if m:
            # code inlined from headers keeps the last line of the kernel's own file; common.cuh is
            # reported separately (negative line numbers)
            base = os.path.basename(m.group(1))
            if base == cubin_stem + ".cu":
                line = int(m.group(2))
            elif base == "common.cuh":
                line = -int(m.group(2))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((line, m.group(2).strip()))
    return out


def main():
    rep, kre, stem = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    # a report may hold several launches of the kernel: keep the first
    first = data[0][0]
    for k in range(1, len(data)):
        if data[k][0] == first:
            data = data[:k]
            break
    sl = sass_lines(stem, sys.argv[5] if len(sys.argv) > 5 else kre)
    if len(sl) != len(data):
        print(f"# WARNING: {len(data)} profiled SASS rows vs {len(sl)} in the current build: the .so differs from the capture")
    si, ii = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    per = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
    for (line, _), r in zip(sl, data):
        p = per[line]
        p[0] += num(r[si])
        p[1] += num(r[ii])
        for c in stall_cols:
            p[2][hdr[c][6:]] += num(r[c])
    tot_s = sum(p[0] for p in per.values()) or 1.0
    tot_i = sum(p[1] for p in per.values()) or 1.0
    src = open(os.path.join(ROOT, "domain_rag_system_b200", "csrc", stem + ".cu")).read().splitlines()
    print(f"# {kre}: {tot_s:.0f} samples, {tot_i / 1e6:.0f} M warp instructions; per source line of {stem}.cu")
    print("# samples%  instr%   line  top stalls                              source")
    for line, p in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        st = sorted(p[2].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{k}:{100 * v / max(p[0], 1):.0f}" for k, v in st)
        text = src[line - 1].strip()[:70] if 0 < line <= len(src) else (f"common.cuh:{-line}" if line < 0 else "")
        print(f"  {100 * p[0] / tot_s:6.2f}  {100 * p[1] / tot_i:6.2f}  {line:5d}  {sts:38s}  {text}")


if __name__ == "__main__":
    main()
