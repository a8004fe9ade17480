#!/usr/bin/env python3
"""Per-source-line instruction counts / stall samples of one kernel out of a .ncu-rep (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
hdr, acc = None, []
for r in csv.reader(io.StringIO(out)):
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[2] != "-":
        continue
    try:
        ie, smp = int(r[hdr.index("Instructions Executed")] or 0), int(r[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    acc.append((ie, smp, r[0], r[1]))
tot_i = sum(a[0] for a in acc) or 1
tot_s = sum(a[1] for a in acc) or 1
print(f"total warp-instr {tot_i}, samples {tot_s}")
for ie, smp, ln, src in sorted(acc, key=lambda a: -a[0])[:top]:
    print(f"{100.0 * ie / tot_i:6.2f}% inst {100.0 * smp / tot_s:6.2f}% smp  L{ln:>5s}  {src.strip()[:120]}")
