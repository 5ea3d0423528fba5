#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 1200 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline --skip-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
ncu --set full --clock-control none --import-source on -k regex:blur_tma -c 1 -f -o gpurun_out/prof_blur_tma python scripts/ncu_misc.py > gpurun_out/ncu_blur2.log 2>&1; echo "ncu rc=$?"
