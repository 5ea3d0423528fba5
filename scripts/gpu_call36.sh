#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python scripts/microbench.py rgb 2>&1 | grep -E "^fwd|^wgrad" | tee gpurun_out/microbench_c36.txt
timeout 500 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline --skip-roofline 2>gpurun_out/bench_c36.err | tee gpurun_out/bench_c36.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
tail -2 gpurun_out/bench_c36.err
