#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "halo or epilogue_bwd or upfirdn2d" 2>&1 | tail -25
python scripts/microbench.py 2>&1 | grep -E "fwd   conv3x3 s1 (32|64)|epilogue|skip|torgb" | tee gpurun_out/microbench_c9.txt
