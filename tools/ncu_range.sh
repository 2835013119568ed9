#!/bin/bash
# ncu --set full of k_range (count measures, hub-heavy sources) at IHub: raw metrics of every
# launch (the passes of the pruned candidate buffer) and the hot lines of the longest one.
set -u
mkdir -p gpurun_out /tmp/ncu
WL=${1:-rmat21}
ncu --set full --clock-control none --import-source on -k regex:"^k_range$" --launch-count 6 \
    -o /tmp/ncu/range_$WL python tools/profile_one.py $WL 0 CN 1 1 > gpurun_out/ncu_range_$WL.log 2>&1
ncu -i /tmp/ncu/range_$WL.ncu-rep --page raw --csv > gpurun_out/r02c_ncu_range_${WL}_raw.csv 2> /dev/null
LONGEST=$(python - <<P
import csv
rows = list(csv.reader(open("gpurun_out/r02c_ncu_range_${WL}_raw.csv")))
hdr = rows[0]; c = hdr.index("gpu__time_duration.sum")
vals = [float(r[c].replace(",", "")) for r in rows[2:]]
print(max(range(len(vals)), key=lambda i: vals[i]))
P
)
echo "longest launch index: $LONGEST" >> gpurun_out/ncu_range_$WL.log
python tools/ncu_hotlines.py /tmp/ncu/range_$WL.ncu-rep "k_range" 40 $LONGEST > gpurun_out/r02c_ncu_range_${WL}_hot.md 2> /dev/null
ls -la /tmp/ncu | tail -3
