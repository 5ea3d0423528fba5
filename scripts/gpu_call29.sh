#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/microbench.py conv 2>&1 | grep -E "fwd|wgrad" | grep -E "torgb|fromrgb|@32|@8|up2 512" | tee gpurun_out/microbench_c29.txt
timeout 600 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline 2>gpurun_out/bench_c29.err | tee gpurun_out/bench_c29.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); [print(l['kernel'], round(l['kernel_ms'],3), round(l['frac'],3)) for l in d['roofline_layers']]"
tail -3 gpurun_out/bench_c29.err
k=c256
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_fwd --launch-skip 1 -c 1 -f -o gpurun_out/r01c_$k python scripts/ncu_shapes.py $k > gpurun_out/ncu_c_$k.log 2>&1; echo "$k rc=$?"
ncu -i gpurun_out/r01c_$k.ncu-rep --page raw --csv > gpurun_out/r01c_${k}_raw.csv 2>/dev/null
rm -f gpurun_out/r01c_$k.ncu-rep
