"""Top stall lines of one kernel from `ncu --page source --csv` (SASS view): python scripts/ncu_hot.py <csv> [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print("total samples", tot)
agg = {c: sum(int(r[idx[c]] or 0) for r in data) for c in stall_cols}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]] or 0))[:n]
for i in sorted(top):
    r = data[i]
    st = {c[6:]: int(r[idx[c]] or 0) for c in stall_cols if int(r[idx[c]] or 0)}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(r[idx['# Samples']]):7d} exec {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:70]:70s} {st}")
