"""Turns the ncu outputs brought back in gpurun_out/ into small tracked summaries under profiles/.
  python scripts/summarise_profiles.py <round-tag> <launches.csv> <full.ncu-rep> [bench.json] [per_launch.json]"""
import collections, csv, json, os, re, subprocess, sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
bench = sys.argv[4] if len(sys.argv) > 4 else None
perl = sys.argv[5] if len(sys.argv) > 5 else None
os.makedirs("profiles", exist_ok=True)
out = [f"# ncu summary {tag}", "", f"Source files (scratch, not tracked): `{launches}`, `{rep}`.",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python scripts/profile_step.py`",
       "(one eager train step of the bench workload: RN-50, 640x960, B=32; per-launch times are cold-cache and serialised:",
       "compare SHARES, not absolutes).", ""]
lines = [l for l in open(launches) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e6 if row["Metric Unit"] == "ns" else (v / 1e3 if row["Metric Unit"] == "us" else v)
    agg[name][0] += 1; agg[name][1] += v; tot += v
out += [f"## Launch list: {sum(a[0] for a in agg.values())} launches, {tot:.2f} ms total", "",
        "| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k[:80]}` | {n} | {ms:.3f} | {ms / tot:.1%} |")
data, hdr, units = [], None, None
for one in rep.split(","):          # several reports may be given, comma separated
    raw = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) > 2:
        hdr, units = rows[0], rows[1]
        data += rows[2:]
if data:
    idx = {h: i for i, h in enumerate(hdr)}
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    out += ["", "## `ncu --set full --clock-control none --import-source on` captures", "",
            "| # | kernel | grid | " + " | ".join(k.split(".")[0].replace("__", " ") for k in keys) + " |",
            "|---|---|---|" + "---:|" * len(keys)]
    for n, d in enumerate(data):
        vals = [f"{d[idx[k]]} {units[idx[k]]}" if k in idx else "-" for k in keys]
        out.append(f"| {n} | `{d[idx['Kernel Name']][:40]}` | {d[idx['Grid Size']]} | " + " | ".join(vals) + " |")
if bench and os.path.exists(bench):
    b = json.load(open(bench))
    out += ["", "## bench.py line of the same build (not under a profiler)", "", "```json", json.dumps(b, indent=1)[:6000], "```"]
if perl and os.path.exists(perl):
    p = json.load(open(perl))
    out += ["", "## per-launch CUDA-event table (Engine.profile_ops, eager, min of 3)", "",
            "| kind | layer | ms | TFLOP/s | GB/s (algorithmic) |", "|---|---|---:|---:|---:|"]
    for r in sorted(p["per_launch"], key=lambda r: -r["ms"])[:60]:
        out.append(f"| {r['kind']} | {r['name']} | {r['ms']:.3f} | {r['flops'] / r['ms'] / 1e9:.0f} | {r['bytes'] / r['ms'] / 1e6:.0f} |")
open(f"profiles/{tag}.md", "w").write("\n".join(out) + "\n")
print("wrote", f"profiles/{tag}.md", len(out), "lines")
