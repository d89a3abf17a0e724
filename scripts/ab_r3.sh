#!/bin/bash
# round-2 (session 3) A/B: tile zigzag + L2 hints, all on one box.  Output under gpurun_out/r3a_*.
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; shift; timeout 300 $B "$@" --profile-json gpurun_out/r3a_prof_$tag.json > gpurun_out/r3a_$tag.json 2> gpurun_out/r3a_$tag.err; echo "$tag: $(python -c "import json,sys; d=json.loads(open('gpurun_out/r3a_$tag.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))" 2>&1 | tail -1)"; }
run base
run zz1 --zigzag 1
run zz3 --zigzag 3
run zz3h5 --zigzag 3 --l2-hints 5
run zz3h15 --zigzag 3 --l2-hints 15
run h15 --l2-hints 15
run zz1h3 --zigzag 1 --l2-hints 3
run base2
