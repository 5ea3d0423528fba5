#!/bin/bash
# kernel tests for the changed FIR / epilogue kernels, microbench, then ncu of the halo conv (tight timeouts)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python scripts/microbench.py blur 2>&1 | grep -E "blur|skip" | tee gpurun_out/microbench_c22.txt
timeout 300 python scripts/microbench.py epilogue 2>&1 | grep -E "epilogue|bias_act|reduce" | tee -a gpurun_out/microbench_c22.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_halo --launch-skip 2 -c 1 -f -o gpurun_out/r01_halo_c32 python scripts/ncu_conv.py 0 c32 > gpurun_out/ncu_halo_c32.log 2>&1; echo "halo c32 rc=$?"
ncu -i gpurun_out/r01_halo_c32.ncu-rep --page raw --csv > gpurun_out/r01_halo_c32_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_halo_c32.ncu-rep --page source --csv > gpurun_out/r01_halo_c32_source.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_halo --launch-skip 2 -c 1 -f -o gpurun_out/r01_halo_c64 python scripts/ncu_conv.py 0 c64 > gpurun_out/ncu_halo_c64.log 2>&1; echo "halo c64 rc=$?"
ncu -i gpurun_out/r01_halo_c64.ncu-rep --page raw --csv > gpurun_out/r01_halo_c64_raw.csv 2>/dev/null
ncu -i gpurun_out/r01_halo_c64.ncu-rep --page source --csv > gpurun_out/r01_halo_c64_source.csv 2>/dev/null
timeout 900 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline --skip-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
