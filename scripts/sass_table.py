"""Per-kernel counts of the Blackwell-specific SASS mnemonics of liburso_b200.so (markdown table on stdout):
  python scripts/sass_table.py [path/to/liburso_b200.so]"""
import collections, re, subprocess, sys

so = sys.argv[1] if len(sys.argv) > 1 else "ursonet_b200/liburso_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKRED", "UTCBAR", "SYNCS", "ACQBULK|PREEXIT", "LDGSTS", "REDG"]
rows, cur, k = collections.OrderedDict(), None, -1
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        full = names[k].replace("void ", "").replace("(int)", "")
        cur = full[:full.index(">(") + 1] if ">(" in full else re.sub(r"\(.*", "", full)
        rows[cur] = collections.Counter()
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
        rows[cur]["n"] += 1
        for c in cols:
            if re.search(r"\b(" + c + r")\b|\b(" + c + r")\.", line):
                rows[cur][c] += 1
print("| kernel | " + " | ".join(c.replace("|", "/") for c in cols) + " | instructions |")
print("|---|" + "---:|" * (len(cols) + 1))
for name in sorted(rows):
    r = rows[name]
    if not name.startswith("urso::"):
        continue
    print(f"| `{name}` | " + " | ".join(str(r[c]) for c in cols) + f" | {r['n']} |")
