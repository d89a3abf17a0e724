"""Bottleneck isolation for Engine F: per-layer forward times with the TMA loads or the MMAs switched off
(URSO_DBG_NO_TMA=1 / URSO_DBG_NO_MMA=1; outputs are garbage, timing only).  Needs a library built with
`make -C ursonet_b200/csrc clean all DEBUG_KNOBS=1`.  Run under gpurun.  The knobs also exist in the CTA-pair kernel:
`URSO_CTA2=1 python scripts/bench_isolate.py` isolates the cta_group::2 path (first measurement planned for round 2)."""
import json, os, subprocess, sys
out = {}
for tag, env in (("normal", {}), ("no_tma", {"URSO_DBG_NO_TMA": "1"}), ("no_mma", {"URSO_DBG_NO_MMA": "1"})):
    e = dict(os.environ); e.update(env)
    pj = f"gpurun_out/iso_{tag}.json"
    subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--profile-json", pj],
                   env=e, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out[tag] = {(r["kind"], r["name"]): r["ms"] for r in json.load(open(pj))["per_launch"] if r["kind"] in ("conv_fwd", "conv_dgrad")}
seen = set()
for k in out["normal"]:
    key = (k[0], k[1][:4] + k[1][5:])
    if key in seen:
        continue
    seen.add(key)
    print(f"{k[0]:10s} {k[1]:22s} normal {out['normal'][k]*1e3:7.1f}  no_tma {out['no_tma'][k]*1e3:7.1f}  no_mma {out['no_mma'][k]*1e3:7.1f}")
