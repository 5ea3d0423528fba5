#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --skip-cpu-baseline --skip-roofline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"
tail -5 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json | cut -c1-600
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-900
