#!/bin/bash
# ncu --set full of the ingest kernels (one launch each) on R-MAT 18 written as a Matrix Market file.
set -u
mkdir -p gpurun_out /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:"k_mtx_|k_ing_" --launch-count 8 \
    -o /tmp/ncu/ingest python tools/ingest_bench.py 18 16 > gpurun_out/ncu_ingest.log 2>&1
ncu -i /tmp/ncu/ingest.ncu-rep --page raw --csv > gpurun_out/r02c_ncu_ingest_raw.csv 2> /dev/null
python tools/ncu_hotlines.py /tmp/ncu/ingest.ncu-rep "k_mtx_parse" 12 > gpurun_out/r02c_ncu_ingest_hot_k_mtx_parse.md 2> /dev/null
tail -2 gpurun_out/ncu_ingest.log
