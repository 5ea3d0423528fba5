#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "packed or halo" 2>&1 | tail -8
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/microbench.py packed 2>&1 | grep -E "fwd|wgrad" | tee gpurun_out/microbench_c26.txt
timeout 600 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline --skip-roofline 2>gpurun_out/bench_c26.err | tee gpurun_out/bench_c26.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
tail -5 gpurun_out/bench_c26.err
