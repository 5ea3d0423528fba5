#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "upfirdn2d or blur" 2>&1 | tail -3
python scripts/microbench.py blur 2>&1 | grep -E "blur" | tee gpurun_out/microbench_c19.txt
timeout 1200 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline --skip-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
