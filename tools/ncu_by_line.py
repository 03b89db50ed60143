#!/usr/bin/env python
"""Aggregate an ncu SASS-level source page by CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <cubin> <mangled kernel name> [top] [demangled-name substring]
Joins `ncu --page source --print-source sass --csv` (per-instruction counters) with
`nvdisasm --print-line-info` (address -> file:line) and prints, per source line: executed warp
instructions, shared-memory wavefronts (total / excessive) and stall samples."""
import csv
import re
import subprocess
import sys
from collections import defaultdict


import os
SORTKEY = "Instructions Executed" if os.environ.get("BY_INST") else "# Samples"


def main():
    rep, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    addr2line, cur, active = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            active = kernel in ln
        if not active:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
        if m and cur:
            addr2line[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    want = sys.argv[5] if len(sys.argv) > 5 else None      # substring of the demangled kernel name
    hdr_i, skip = None, int(os.environ.get("KERNEL_INDEX", "0"))      # KERNEL_INDEX=n: the n-th matching launch
    for i, r in enumerate(rows):
        if r and r[0] == "Kernel Name" and (want is None or want in r[1]):
            if skip > 0:
                skip -= 1
                continue
            hdr_i = next(j for j in range(i, len(rows)) if rows[j] and rows[j][0] == "Address")
            break
    hdr = rows[hdr_i]
    col = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared",
                                     "L1 Wavefronts Shared Excessive", "stall_long_sb", "stall_barrier", "stall_short_sb",
                                     "stall_wait", "stall_mio", "stall_math", "L2 Theoretical Sectors Global")}
    agg = defaultdict(lambda: defaultdict(float))
    first = None
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or r[0] in ("Address", "Kernel Name"):
            break                                   # only the first matching kernel instance
        try:
            a = int(r[col["Address"]], 16)
        except ValueError:
            continue
        if first is None:
            first = a
        key = addr2line.get(a - first, ("?", 0))
        for n, c in col.items():
            if n in ("Address", "Source"):
                continue
            try:
                agg[key][n] += float(r[c] or 0)
            except ValueError:
                pass
    tot = defaultdict(float)
    for k, v in agg.items():
        for n, x in v.items():
            tot[n] += x
    print("TOTAL inst %.3g  smem wavefronts %.3g (excessive %.3g)  samples %.0f" %
          (tot["Instructions Executed"], tot["L1 Wavefronts Shared"], tot["L1 Wavefronts Shared Excessive"], tot["# Samples"]))
    print("%-26s %8s %6s %9s %9s %7s | %6s %6s %6s %6s %6s %6s" % ("line", "inst", "inst%", "smem_wf", "smem_exc", "smpl%",
                                                                 "longsb", "barr", "shortsb", "wait", "mio", "math"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][SORTKEY])[:top]:
        print("%-26s %8.3g %5.1f%% %9.3g %9.3g %6.1f%% | %6.0f %6.0f %6.0f %6.0f %6.0f %6.0f" % (
            "%s:%d" % k, v["Instructions Executed"], 100 * v["Instructions Executed"] / max(tot["Instructions Executed"], 1),
            v["L1 Wavefronts Shared"], v["L1 Wavefronts Shared Excessive"], 100 * v["# Samples"] / max(tot["# Samples"], 1),
            v["stall_long_sb"], v["stall_barrier"], v["stall_short_sb"], v["stall_wait"], v["stall_mio"], v["stall_math"]))


if __name__ == "__main__":
    main()
