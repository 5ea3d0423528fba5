#!/bin/bash
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 1200 python bench.py --steps 16 --warmup 3 --skip-cpu-baseline > gpurun_out/bench_r1_graph.json 2> gpurun_out/bench_r1_graph.err
tail -5 gpurun_out/bench_r1_graph.err; cat gpurun_out/bench_r1_graph.json
timeout 600 python bench.py --steps 8 --warmup 3 --skip-cpu-baseline --no-graph --skip-roofline 2>&1 | tail -2
python scripts/profile_step.py --rows 30 > gpurun_out/prof_plain2.txt 2>&1; grep -v "^-" gpurun_out/prof_plain2.txt | cut -c1-72,130-230 | head -40
