#!/bin/bash
# round-1 record: GPU tests, smoke, full bench line (with roofline + cpu_baseline), reference arm, per-kernel step profile,
# ncu launch list of the bench command
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 16 --warmup 3 2>gpurun_out/bench_c38.err | tee gpurun_out/bench_c38.json | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 2>/dev/null | tee gpurun_out/bench_ref_c38.json | cut -c1-200
timeout 300 python scripts/profile_step.py --rows 60 > gpurun_out/profile_plain_c38.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c38.csv python bench.py --steps 1 --warmup 3 --no-reg --ncu-range --skip-cpu-baseline --skip-roofline > gpurun_out/ncu_bench_c38.log 2>&1; echo "ncu rc=$?"
tail -1 gpurun_out/launches_c38.csv | cut -c1-200
