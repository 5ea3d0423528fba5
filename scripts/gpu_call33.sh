#!/bin/bash
timeout 100 python scripts/microbench.py packed > gpurun_out/microbench_c33.txt 2>&1; echo "rc=$?"; grep -E "^fwd|^wgrad|Error|error" gpurun_out/microbench_c33.txt | head -12
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "packed or halo" 2>&1 | tail -3
