#!/bin/bash
for k in c32 c64; do
timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_fwd --launch-skip 1 -c 1 -f -o gpurun_out/r01d_$k python scripts/ncu_shapes.py $k > gpurun_out/ncu_d_$k.log 2>&1; echo "$k rc=$?"
ncu -i gpurun_out/r01d_$k.ncu-rep --page raw --csv > gpurun_out/r01d_${k}_raw.csv 2>/dev/null
ncu -i gpurun_out/r01d_$k.ncu-rep --page source --csv > gpurun_out/r01d_${k}_source.csv 2>/dev/null
rm -f gpurun_out/r01d_$k.ncu-rep
done
