#!/bin/bash
# tests, then a first bench at reduced steps, then the launch list of one step
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r1_simt.json 2> gpurun_out/bench_r1_simt.err
tail -3 gpurun_out/bench_r1_simt.err; cat gpurun_out/bench_r1_simt.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r1_simt.csv \
  python bench.py --steps 1 --warmup 1 --no-reg --skip-cpu-baseline --skip-roofline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log
wc -l gpurun_out/launches_r1_simt.csv
