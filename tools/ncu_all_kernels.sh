#!/bin/bash
# One `ncu --set full` launch per kernel of interest (round 2 evidence, profiles/r02_ncu_*.md):
# source-centric kernels on IHub / LHub R-MAT 18 (BASELINE configs[0]), bucket path + top-K on R-MAT 22 D=16.
# Run on a GPU box AFTER the same commands have exited 0 without ncu.  Reports land in gpurun_out/.
set -u
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on"
# IHub count + float: frontier, tiny, hash, range, range_flt, dense, general top-K
ncu $COMMON -k regex:"k_work_short|k_bin|k_tiny|k_hash|k_range|k_dense|k_select_hist|k_select_compact|k_flt_item" \
    --launch-skip 0 --launch-count 40 -o gpurun_out/r02_ihub18 python tools/profile_one.py rmat18 0 CN,AA 1 1 > gpurun_out/ncu_ihub18.log 2>&1
# LHub on the source path (k_elig, k_work_short<true> with compaction)
ncu $COMMON -k regex:"k_elig|k_work_short|k_work_long" --launch-count 6 -o gpurun_out/r02_lhub18_source \
    python tools/profile_one.py rmat18 16 JC 1 1 > gpurun_out/ncu_lhub18.log 2>&1
# bucket path, second round of predictions (plan resident): k_bucket, k_score, select, ordered compaction, sort passes, detour
ncu $COMMON -k regex:"k_bucket|k_score|k_sel11_hist|k_sel11_step|k_ordered_count2|k_ordered_write2|k_scatter|k_tilehist|k_rowscan|k_pair_emit|k_pair_reduce|k_big_place" \
    --launch-skip 84 --launch-count 45 -o gpurun_out/r02_bucket22 python tools/profile_one.py rmat22 16 JC,AA 2 > gpurun_out/ncu_bucket22.log 2>&1
# plan build + graph preparation + batch kernels (first prediction on a new graph)
ncu $COMMON -k regex:"k_plan|k_validate|k_symmetry|k_del_|k_degrees" --launch-count 30 -o gpurun_out/r02_plan22 \
    python tools/profile_one.py rmat22 16 JC 1 > gpurun_out/ncu_plan22.log 2>&1
ls -la gpurun_out/*.ncu-rep
