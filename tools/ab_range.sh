#!/bin/bash
# A/B of the k_range knobs at IHub (common neighbours): bash tools/ab_range.sh [workload] [variants...]
WL=${1:-rmat22}; shift
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --workload $WL --degree 0 --measures CN --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/$tag.json 2> gpurun_out/$tag.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/$tag.json').read().strip().splitlines()[-1]); print('$tag', d['ms_per_step'], d['bins'])"; }
for v in "${@:-base}"; do
  case $v in
    base) run base X=1 ;;
    nofence) run nofence NLP_B200_RANGE_FENCE=0 ;;
    noquarter) run noquarter NLP_B200_RANGE_QUARTER=0 ;;
    neither) run neither NLP_B200_RANGE_FENCE=0 NLP_B200_RANGE_QUARTER=0 ;;
  esac
done
