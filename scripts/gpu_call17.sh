#!/bin/bash
python scripts/profile_step.py --rows 44 > gpurun_out/prof_plain5.txt 2>&1; grep -v "^-" gpurun_out/prof_plain5.txt | grep -v "autograd::engine" | cut -c1-72,130-230 | head -46
