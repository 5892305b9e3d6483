#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS listing of one kernel with `nvdisasm --print-line-info` of the same
cubin, and aggregate executed warp instructions / stall samples per source line.

  ncu -i rep.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all build/csrc/holo_realize.o ; nvdisasm --print-line-info -c X.cubin > dis.txt
  python profiles/sass_by_line.py src.csv dis.txt '<mangled kernel name>' [top]
"""
import csv, re, sys, collections

src, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
head = rows[hdr]
ci, cs, csamp = head.index("Instructions Executed"), head.index("Source"), head.index("# Samples")
ins = []
for r in rows[hdr + 1:]:
    if len(r) <= ci or not r[0].startswith("0x"):
        continue
    ins.append((int(r[0], 16), r[cs].strip(), int(r[ci] or 0), int(r[csamp] or 0)))
base = ins[0][0]
# disassembly: offsets -> (file, line)
lines = open(dis).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith("\t.section\t.text." + kern)][0]
loc = {}
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        loc[int(m.group(1), 16)] = (cur, m.group(2).strip())
agg = collections.Counter()
samp = collections.Counter()
miss = 0
tot = sum(x[2] for x in ins)
tots = sum(x[3] for x in ins)
for addr, text, n, s in ins:
    ent = loc.get(addr - base)
    if ent is None:
        miss += 1
        key = ("?", 0)
    else:
        key = ent[0]
    agg[key] += n
    samp[key] += s
print(f"# {kern}: {tot:.4g} warp instructions, {tots} stall samples, {miss} unmatched SASS rows")
print(f"{'file:line':32s} {'inst %':>7s} {'samples %':>9s}")
for key, n in agg.most_common(top):
    print(f"{key[0] + ':' + str(key[1]):32s} {100.0 * n / tot:7.2f} {100.0 * samp[key] / max(tots, 1):9.2f}")
if "--dump" in sys.argv:
    import json
    json.dump({f"{k[0]}:{k[1]}": [agg[k], samp[k]] for k in agg}, open("/tmp/by_line.json", "w"))
