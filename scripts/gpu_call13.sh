#!/bin/bash
python scripts/profile_step.py --rows 40 > gpurun_out/prof_plain4.txt 2>&1; grep -v "^-" gpurun_out/prof_plain4.txt | cut -c1-72,130-230 | head -48
ncu --set full --clock-control none --import-source on -k regex:blur_rows -c 1 -f -o gpurun_out/prof_blur python scripts/ncu_misc.py > gpurun_out/ncu_blur.log 2>&1; echo "ncu blur rc=$?"
