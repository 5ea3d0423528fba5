#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "umma" 2>&1 | tail -30
