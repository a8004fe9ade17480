#!/usr/bin/env python3
"""Per-source-line instruction counts / stall samples of one kernel out of a .ncu-rep.
ncu's CSV source page is SASS-only, so the SASS rows are joined (by instruction order) with `nvdisasm -g` of the
cubin inside the shipped library, which carries the `//## File ..., line N` markers of -lineinfo.
usage: tools/ncu_lines.py report.ncu-rep kernel_regex lib.so [top_n]"""
import csv, io, os, re, subprocess, sys, tempfile

rep, kre, lib = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the page may hold several launches of the kernel back to back: keep the first block
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
body = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
kname = rows[start - 1][1].split("(")[0] if start > 0 else kre
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
ct = hdr.index("Thread Instructions Executed")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    base = re.sub(r"<.*", "", kname).split()[-1]
    for m in re.finditer(r"^\.text\.(\S*%s\S*):$" % re.escape(base), txt, re.M):   # template instances: take the one of equal length
        seg = txt[m.end():]
        nxt = re.search(r"^//-+ \.text\.", seg, re.M)
        seg = seg[: nxt.start()] if nxt else seg
        cur, cand = 0, []
        for ln in seg.splitlines():
            mm = re.search(r"//## File \"([^\"]+)\", line (\d+)", ln)
            if mm:
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                cand.append(cur)
        if len(cand) == len(body):
            lines = cand
            break
    if lines is not None:
        break
if lines is None or len(lines) != len(body):
    sys.exit(f"cannot align: {0 if lines is None else len(lines)} disassembled vs {len(body)} profiled instructions")
agg = {}
tot_i = tot_s = 0
for r, key in zip(body, lines):
    n, s, t = int(r[ci] or 0), int(r[cs] or 0), int(r[ct] or 0)
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += n; a[1] += s; a[2] += t
    tot_i += n; tot_s += s
src_cache = {}
def src(key):
    if not key:
        return ""
    f, l = key
    if f not in src_cache:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return src_cache[f][l - 1].strip()[:110] if 0 < l <= len(src_cache[f]) else ""
print(f"{kname}: {tot_i} warp instructions, {tot_s} samples, {len(body)} SASS instructions")
for key, (n, s, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100.0 * n / tot_i:5.2f}% inst {100.0 * s / max(tot_s, 1):5.2f}% samp  thr/inst {t / max(n, 1):4.1f}  {key[0] if key else '?'}:{key[1] if key else 0}: {src(key)}")
