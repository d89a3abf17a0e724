"""Dynamic warp-instruction counts and stall samples per CUDA source line of one kernel: joins `ncu --page source --csv`
(SASS rows, address order) with `nvdisasm -g` line info of the same build.
  python scripts/ncu_lines.py <ncu_sass.csv> <nvdisasm.sass> <mangled-substring> <divide-by>"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
key = sys.argv[3]; div = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
insec = False; cur = None; locs = []
for l in open(sys.argv[2]):
    if l.startswith(".text."):
        insec = key in l; continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,6}\*/', l):
        locs.append(cur)
print("sass rows", len(data), "disasm instrs", len(locs))
n = min(len(data), len(locs))
ex = collections.Counter(); sm = collections.Counter()
for r, loc in zip(data[:n], locs[:n]):
    ex[loc] += int(r[idx["Instructions Executed"]] or 0); sm[loc] += int(r[idx["# Samples"]] or 0)
tot = sum(ex.values()); print("total warp instrs", tot, "per unit", tot / div)
for loc, v in sorted(ex.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{loc[0]:28s} {loc[1]:5d}  exec/unit {v / div:9.1f}  samples {sm[loc]:6d}")
