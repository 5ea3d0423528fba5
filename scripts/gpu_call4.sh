#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "umma" 2>&1 | tail -30
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 1200 python bench.py --steps 8 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_r1_umma.json 2> gpurun_out/bench_r1_umma.err
tail -3 gpurun_out/bench_r1_umma.err; cat gpurun_out/bench_r1_umma.json
