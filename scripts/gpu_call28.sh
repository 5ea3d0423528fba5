#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "packed or halo" 2>&1 | tail -5
timeout 300 python scripts/microbench.py conv 2>&1 | grep -E "fwd" | tee gpurun_out/microbench_c28.txt
for k in c128; do
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_fwd --launch-skip 1 -c 1 -f -o gpurun_out/r01c_$k python scripts/ncu_shapes.py $k > gpurun_out/ncu_c_$k.log 2>&1; echo "$k rc=$?"
ncu -i gpurun_out/r01c_$k.ncu-rep --page raw --csv > gpurun_out/r01c_${k}_raw.csv 2>/dev/null
ncu -i gpurun_out/r01c_$k.ncu-rep --page source --csv > gpurun_out/r01c_${k}_source.csv 2>/dev/null
rm -f gpurun_out/r01c_$k.ncu-rep
done
k=c32w
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad --launch-skip 1 -c 1 -f -o gpurun_out/r01c_$k python scripts/ncu_shapes.py c32 wgrad > gpurun_out/ncu_c_$k.log 2>&1; echo "$k rc=$?"
ncu -i gpurun_out/r01c_$k.ncu-rep --page raw --csv > gpurun_out/r01c_${k}_raw.csv 2>/dev/null
ncu -i gpurun_out/r01c_$k.ncu-rep --page source --csv > gpurun_out/r01c_${k}_source.csv 2>/dev/null
rm -f gpurun_out/r01c_$k.ncu-rep
