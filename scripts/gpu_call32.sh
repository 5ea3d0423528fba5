#!/bin/bash
timeout 100 python scripts/microbench.py packed > gpurun_out/microbench_c32.txt 2>&1; echo "rc=$?"; tail -12 gpurun_out/microbench_c32.txt
