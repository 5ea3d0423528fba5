#!/bin/bash
timeout 300 python scripts/profile_step.py --rows 60 > gpurun_out/profile_plain_c27.txt 2>&1; echo "plain rc=$?"
timeout 300 python scripts/profile_step.py --reg --rows 60 > gpurun_out/profile_reg_c27.txt 2>&1; echo "reg rc=$?"
for k in upfused c32; do
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_fwd --launch-skip 1 -c 1 -f -o gpurun_out/r01c_$k python scripts/ncu_shapes.py $k > gpurun_out/ncu_c_$k.log 2>&1; echo "$k rc=$?"
ncu -i gpurun_out/r01c_$k.ncu-rep --page raw --csv > gpurun_out/r01c_${k}_raw.csv 2>/dev/null
ncu -i gpurun_out/r01c_$k.ncu-rep --page source --csv > gpurun_out/r01c_${k}_source.csv 2>/dev/null
rm -f gpurun_out/r01c_$k.ncu-rep
done
