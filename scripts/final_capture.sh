#!/bin/bash
# End-of-round capture on one box: full GPU test suite, bench line, ncu launch list, DRAM traffic of the conv engines,
# full ncu captures of representative launches, bench lines of the other BASELINE configs.  Output: gpurun_out/<tag>_*
T=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/${T}_prof.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv python scripts/profile_step.py > gpurun_out/${T}_ncu1.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k "regex:conv_gemm|wgrad" --csv --log-file gpurun_out/${T}_traffic.csv python scripts/profile_step.py > gpurun_out/${T}_ncu2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${T}_layers python scripts/profile_layers.py conv_fwd:conv1 conv_fwd:res2a_branch2b conv_fwd:res2a_branch2c conv_dgrad:res2a_out conv_dgrad:res2a_branch2a conv_fwd:res3a_branch2b conv_fwd:res4b_branch2b conv_fwd:res4b_branch2c conv_wgrad:res2a_branch2b conv_wgrad:res4b_branch2b conv_wgrad:conv1 conv_fwd:res4b_branch2a > gpurun_out/${T}_ncu3.log 2>&1; tail -2 gpurun_out/${T}_ncu3.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --regress_ori --width 1920 --height 1200 --batch 16 > gpurun_out/${T}_cfg3.json 2> gpurun_out/${T}_cfg3.err; tail -c 300 gpurun_out/${T}_cfg3.json | head -c 300; echo
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --backbone resnet101 --ori_resolution 24 --batch 8 > gpurun_out/${T}_cfg4.json 2> gpurun_out/${T}_cfg4.err
python - <<PY
import json
for t in ("bench","cfg3","cfg4"):
    try:
        d=json.loads(open("gpurun_out/${T}_%s.json"%t).read().strip().splitlines()[-1]); print(t, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "frac", d["roofline"]["frac"])
    except Exception as e: print(t, "failed", e)
PY
