#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python scripts/microbench.py blur 2>&1 | grep -E "blur|skip" | tee gpurun_out/microbench_c23.txt
timeout 300 python scripts/microbench.py epilogue 2>&1 | grep -E "epilogue|bias_act|reduce" | tee -a gpurun_out/microbench_c23.txt
timeout 900 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline --skip-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
for k in c32 c64; do
timeout 150 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_halo --launch-skip 1 -c 1 -f -o gpurun_out/r01_halo_$k python scripts/ncu_conv.py 0 $k > gpurun_out/ncu_halo_$k.log 2>&1; echo "halo $k rc=$?"
ncu -i gpurun_out/r01_halo_$k.ncu-rep --page raw --csv > gpurun_out/r01_halo_${k}_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_halo_$k.ncu-rep --page source --csv > gpurun_out/r01_halo_${k}_source.csv 2>/dev/null
done
timeout 300 python scripts/profile_step.py --rows 40 > gpurun_out/profile_plain_c23.txt 2>&1
timeout 300 python scripts/profile_step.py --rows 40 --reg > gpurun_out/profile_reg_c23.txt 2>&1
