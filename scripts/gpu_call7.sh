#!/bin/bash
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 1200 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_r1_c7.json 2> gpurun_out/bench_r1_c7.err
tail -5 gpurun_out/bench_r1_c7.err; cat gpurun_out/bench_r1_c7.json
python scripts/profile_step.py --rows 34 > gpurun_out/prof_plain3.txt 2>&1; grep -v "^-" gpurun_out/prof_plain3.txt | cut -c1-72,130-230 | head -44
