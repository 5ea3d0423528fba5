#!/bin/bash
python scripts/profile_step.py > gpurun_out/prof_plain.txt 2>&1; grep -v "^-" gpurun_out/prof_plain.txt | cut -c1-72,130-230 | head -60
