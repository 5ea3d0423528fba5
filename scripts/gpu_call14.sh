#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "upfirdn2d or blur" 2>&1 | tail -5
python scripts/microbench.py blur 2>&1 | grep -E "blur|skip" | tee gpurun_out/microbench_c14.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline --skip-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
