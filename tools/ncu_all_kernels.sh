#!/bin/bash
# One `ncu --set full` launch per kernel of interest (round 2 evidence, profiles/r02_ncu_*.md).
# Run on a GPU box AFTER the same commands have exited 0 without ncu.  The .ncu-rep files stay in
# /tmp (they are far larger than what gpurun copies back); the raw metric tables and the hot
# source lines are exported here, on the box, into gpurun_out/.
set -u
mkdir -p gpurun_out /tmp/ncu
COMMON="--set full --clock-control none --import-source on"
exp() {   # exp <report> <kernel regex for hot lines> ...
  rep=$1; shift
  ncu -i /tmp/ncu/$rep.ncu-rep --page raw --csv > gpurun_out/r02_ncu_${rep}_raw.csv 2> /dev/null
  for k in "$@"; do
    python tools/ncu_hotlines.py /tmp/ncu/$rep.ncu-rep "$k" 14 > gpurun_out/r02_ncu_${rep}_hot_$(echo $k | tr -cd 'a-z0-9_').md 2> /dev/null
  done
}
# A: IHub count + float on BASELINE configs[0] (source-centric kernels, general top-K)
ncu $COMMON -k regex:"k_work_short|k_bin|k_tiny|k_hash|k_range|k_dense|k_select_hist|k_select_compact" --launch-count 28 \
    -o /tmp/ncu/ihub18 python tools/profile_one.py rmat18 0 CN,AA 1 1 > gpurun_out/ncu_ihub18.log 2>&1
exp ihub18 "k_range_flt" "k_range<" "k_hash" "k_tiny"
# B: bucket path + exact top-K + detour on BASELINE configs[1], second prediction (plan resident)
ncu $COMMON -k regex:"k_bucket|k_score|k_sel11_hist|k_ordered_count2|k_ordered_write2|k_scatter|k_tilehist|k_pair_emit|k_pair_reduce|k_big_place" \
    --launch-skip 30 --launch-count 30 -o /tmp/ncu/bucket22 python tools/profile_one.py rmat22 16 JC 2 > gpurun_out/ncu_bucket22.log 2>&1
exp bucket22 "k_bucket" "k_score" "k_ordered_write2"
# C: graph preparation, plan build, batch apply, LHub frontier of the source path
ncu $COMMON -k regex:"k_validate_entries|k_symmetry|k_plan_items|k_plan_scatter|k_del_mark|k_del_write|k_batch_slots" --launch-count 9 \
    -o /tmp/ncu/plan22 python tools/profile_one.py rmat22 16 JC 1 > gpurun_out/ncu_plan22.log 2>&1
exp plan22 "k_validate_entries"
ls -la /tmp/ncu gpurun_out | tail -30
