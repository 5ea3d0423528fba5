#!/bin/bash
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "packed or halo" 2>&1 | tail -4
timeout 200 python scripts/microbench.py conv 2>&1 | grep -E "^fwd" | grep -E "@1024|@512" | tee gpurun_out/microbench_c31.txt
