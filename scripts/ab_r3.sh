#!/bin/bash
# same-box A/B of bench.py variants.  usage: scripts/ab_r3.sh <prefix> "tag:flags" "tag:flags" ...   (output under gpurun_out/<prefix>_*)
mkdir -p gpurun_out
P=$1; shift
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for spec in "$@"; do
  tag=${spec%%:*}; flags=${spec#*:}
  timeout 300 $B $flags --profile-json gpurun_out/${P}_prof_$tag.json > gpurun_out/${P}_$tag.json 2> gpurun_out/${P}_$tag.err
  echo "$tag: $(python -c "import json,sys; d=json.loads(open('gpurun_out/${P}_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']['by_kind']; print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), ' fwd %.3f dgrad %.3f wgrad %.3f' % (r['conv_fwd']['ms'], r['conv_dgrad']['ms'], r['conv_wgrad']['ms']))" 2>&1 | tail -1)"
done
