#!/usr/bin/env python
"""Attribute the warp-stall samples of an `ncu --set full --import-source on` capture to SOURCE lines.

    ncu -i k.ncu-rep --page source --csv > k_source.csv                      (one row per SASS instruction, in order)
    cuobjdump -xelf all build/obj/ba_kernels.o; nvdisasm -g -c *.cubin > k.sass     (the same build: SASS with //## File ... line N)
    python tools/ncu_source_lines.py k_source.csv k.sass <mangled kernel name> <source file> [top N]

The two listings hold the kernel's instructions in the same order; the script joins them by position, sums samples /
executed instructions / stall reasons per source line and prints the top lines (share of samples, share of instructions, the
three largest stall reasons)."""
import csv, re, sys, collections
# usage: map.py <source.csv> <sass file> <function text label> <source file>
csvf, sassf, label, srcf = sys.argv[1:5]
rows = list(csv.reader(open(csvf)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
# sass with line info
lines = open(sassf).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + label))
cur = None; inst_lines = []
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): 
        if inst_lines: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): inst_lines.append((cur, l.strip()))
print("sass insts", len(inst_lines), "csv insts", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
for (loc, txt), r in zip(inst_lines, data):
    key = loc
    a = agg[key]
    a["samples"] += int(r[ix["# Samples"]] or 0)
    a["inst"] += int(r[ix["Instructions Executed"]] or 0)
    for s_ in stalls: a[s_] += int(r[ix[s_]] or 0)
tot = sum(a["samples"] for a in agg.values()); toti = sum(a["inst"] for a in agg.values())
src = open(srcf).read().split("\n")
print("total samples", tot, "inst", toti)
out = []
for key, a in agg.items():
    if key is None: continue
    out.append((key, a))
out.sort(key=lambda x: -x[1]["samples"])
for key, a in out[:int(sys.argv[5]) if len(sys.argv) > 5 else 45]:
    top = sorted(((a[s_], s_) for s_ in stalls), reverse=True)[:3]
    f, ln = key
    stall_txt = ' '.join('%s=%.0f%%' % (s_[6:], 100 * v / max(1, a['samples'])) for v, s_ in top)
    text = src[ln - 1].strip()[:90] if f == srcf.split("/")[-1] and ln <= len(src) else f
    print(f"{f}:{ln:5d} {100*a['samples']/tot:5.1f}% smp {100*a['inst']/toti:5.1f}% inst  {stall_txt} | {text}")
