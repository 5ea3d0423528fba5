#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -k "halo" 2>&1 | grep -E "HALO|passed|failed"
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "umma" 2>&1 | tail -3
python scripts/microbench.py conv 2>&1 | grep -E "conv" | tee gpurun_out/microbench_c11.txt
