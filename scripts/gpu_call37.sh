#!/bin/bash
for k in torgb; do
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pw_small --launch-skip 1 -c 1 -f -o gpurun_out/r01c_$k python scripts/ncu_shapes.py $k > gpurun_out/ncu_c_$k.log 2>&1; echo "$k rc=$?"
ncu -i gpurun_out/r01c_$k.ncu-rep --page raw --csv > gpurun_out/r01c_${k}_raw.csv 2>/dev/null
ncu -i gpurun_out/r01c_$k.ncu-rep --page source --csv > gpurun_out/r01c_${k}_source.csv 2>/dev/null
ncu -i gpurun_out/r01c_$k.ncu-rep --page details > gpurun_out/r01c_${k}_details.txt 2>/dev/null
rm -f gpurun_out/r01c_$k.ncu-rep
done
