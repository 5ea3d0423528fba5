#!/bin/bash
python scripts/microbench.py 2>&1 | grep -v Warning | tee gpurun_out/microbench_c8.txt
