#!/bin/bash
# same-box A/B of two builds of the library.  usage: scripts/ab_lib.sh <prefix> <alt .so> [bench flags]   (output gpurun_out/<prefix>_*)
mkdir -p gpurun_out
P=$1; ALT=$2; shift 2
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline $@"
for tag in alt new alt2 new2; do
  if [ "${tag:0:3}" = "alt" ]; then export URSO_LIB_PATH=$ALT; else unset URSO_LIB_PATH; fi
  timeout 300 $B --profile-json gpurun_out/${P}_prof_$tag.json > gpurun_out/${P}_$tag.json 2> gpurun_out/${P}_$tag.err
  echo "$tag: $(python -c "import json,sys; d=json.loads(open('gpurun_out/${P}_$tag.json').read().strip().splitlines()[-1]); r=d['roofline']['by_kind']; print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), ' fwd %.3f dgrad %.3f wgrad %.3f' % (r['conv_fwd']['ms'], r['conv_dgrad']['ms'], r['conv_wgrad']['ms']))" 2>&1 | tail -1)"
done
