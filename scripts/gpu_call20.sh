#!/bin/bash
# re-entry check: GPU tests, full bench line, per-kernel step profile, ncu launch list of the bench command
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 1200 python bench.py --steps 16 --warmup 3 2>gpurun_out/bench_c20.err | tee gpurun_out/bench_c20.json | cut -c1-600
timeout 600 python scripts/profile_step.py --rows 60 > gpurun_out/profile_plain_c20.txt 2>&1
timeout 600 python scripts/profile_step.py --rows 60 --reg > gpurun_out/profile_reg_c20.txt 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c20.csv python bench.py --steps 1 --warmup 3 --no-reg --ncu-range --skip-cpu-baseline --skip-roofline > gpurun_out/ncu_bench_c20.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/launches_c20.csv | cut -c1-300
