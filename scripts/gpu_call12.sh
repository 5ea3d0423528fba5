#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python scripts/microbench.py conv 2>&1 | grep -E "conv" | tee gpurun_out/microbench_c12.txt
timeout 1200 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_r1_c12.json 2> gpurun_out/bench_r1_c12.err
tail -3 gpurun_out/bench_r1_c12.err; cat gpurun_out/bench_r1_c12.json
