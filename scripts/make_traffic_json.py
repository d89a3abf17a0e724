"""profiles/<tag>_traffic.json from an ncu CSV of one eager train step:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --profile-from-start off -k regex:conv_gemm|wgrad --csv --log-file gpurun_out/traffic.csv python scripts/profile_step.py
  python scripts/make_traffic_json.py r02 gpurun_out/traffic.csv gpurun_out/prof.json"""
import collections, csv, json, sys

tag, path, prof = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(path) if l.startswith('"')]
per = collections.defaultdict(lambda: collections.defaultdict(float))
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    if row["Metric Name"].startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    else:
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[u]       # -> ms
    per[row["ID"]]["name"] = row["Kernel Name"]
    per[row["ID"]][row["Metric Name"]] += v
p = json.load(open(prof))
alg = collections.defaultdict(lambda: [0.0, 0])
for r in p["per_launch"]:
    eng = {"conv_fwd": "conv_gemm", "conv_dgrad": "conv_gemm", "conv_wgrad": "wgrad"}.get(r["kind"])
    if eng:
        alg[eng][0] += r["bytes"]
        alg[eng][1] += r.get("launches", 1)
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                 "-k regex:conv_gemm|wgrad, one eager train step (scripts/profile_step.py), RN-50 640x960 B=32; algorithmic "
                 "bytes = Engine.profile_ops accounting (inputs + outputs + addend + mask bits + weights)", "engines": {}}
for eng in ("conv_gemm", "wgrad"):
    ks = [d for d in per.values() if (eng + "_kernel") in d["name"] and "dense" not in d["name"]]
    n = len(ks)
    dram = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in ks)
    out["engines"][eng] = {"launches": n, "dram_bytes_per_launch": dram / max(n, 1),
                           "algorithmic_bytes_per_launch": alg[eng][0] / max(alg[eng][1], 1),
                           "ratio": dram / max(alg[eng][0], 1.0), "ncu_ms_total": sum(d["gpu__time_duration.sum"] for d in ks)}
json.dump(out, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
