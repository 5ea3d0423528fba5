#!/bin/bash
# final sanity of HEAD: GPU tests, default-config bench, and the per-GPU batch of BASELINE.json configs[2] (4 / GPU)
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 8 --warmup 3 --skip-cpu-baseline --skip-roofline 2>/dev/null | tee gpurun_out/bench_c42.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('b16', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 300 python bench.py --batch 4 --steps 8 --warmup 3 --skip-cpu-baseline --skip-roofline 2>/dev/null | tee gpurun_out/bench_b4_c42.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('b4', d['value'], d['ms_per_step'], d['e2e']['value'])"
