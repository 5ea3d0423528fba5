#!/bin/bash
timeout 100 python scripts/microbench.py packed > gpurun_out/microbench_c34.txt 2>&1; echo "rc=$?"; grep -E "^fwd|^wgrad|Error|error" gpurun_out/microbench_c34.txt | head -12
timeout 100 python scripts/microbench.py conv 2>&1 | grep -E "^fwd" | grep -E "@1024|@512" | head -4
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
