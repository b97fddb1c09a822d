#!/usr/bin/env python
"""Warp-instructions executed per CUDA source line of one kernel: joins `ncu --page source --csv` (per-SASS-instruction
counts of a capture made with --import-source on) with `nvdisasm --print-line-info` of the same kernel in the shipped
libcvpp.so (built with -lineinfo).  Usage: ncu_lines.py report.ncu-rep <kernel-regex> <mangled-name-substring> [topN]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kregex, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "computervision", "pytorch_b200", "libcvpp.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(txt) if l.startswith(".text.") and mangled in l]
    if not starts:
        continue
    lines, cur = [], None
    for l in txt[starts[0] + 1:]:
        if l.startswith(".text.") or l.startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    break
assert lines, "kernel not found in libcvpp.so"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex, "--launch-skip", "0",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
ex = [(int(r[iex]), int(r[ismp] or 0)) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
assert len(ex) == len(lines), (len(ex), len(lines), "the report was captured with a different build of the kernel")
agg, smp = collections.Counter(), collections.Counter()
for (e, s), ln in zip(ex, lines):
    agg[ln] += e
    smp[ln] += s
tot = sum(agg.values())
src = {}
print(f"total warp-instructions {tot}")
for (f, ln), c in agg.most_common(top):
    if f not in src:
        pth = os.path.join(ROOT, "computervision", "pytorch_b200", "csrc", f)
        src[f] = open(pth).read().split("\n") if os.path.exists(pth) else None
    t = src[f][ln - 1].strip()[:100] if src[f] and ln <= len(src[f]) else ""
    print(f"{100 * c / tot:5.1f}% {c:10d} samples {smp[(f, ln)]:5d}  {f}:{ln}  {t}")
