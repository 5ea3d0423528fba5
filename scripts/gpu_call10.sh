#!/bin/bash
python scripts/ncu_conv.py 2 c32 > gpurun_out/plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/plain.log
ncu --set full --clock-control none -k regex:conv_fwd_umma -c 2 -f -o gpurun_out/prof_fwd_v1 python scripts/ncu_conv.py 2 c32 > gpurun_out/ncu1.log 2>&1; echo "ncu1 rc=$?"; tail -5 gpurun_out/ncu1.log
ncu --set full --clock-control none --import-source on -k regex:conv_fwd_umma -c 6 -f -o gpurun_out/prof_fwd_v1b python scripts/ncu_conv.py 2 > gpurun_out/ncu1b.log 2>&1; echo "ncu1b rc=$?"; tail -5 gpurun_out/ncu1b.log
ncu --set full --clock-control none -k regex:conv_fwd_halo -c 2 -f -o gpurun_out/prof_fwd_halo python scripts/ncu_conv.py 0 c32 > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"
ncu --set full --clock-control none -k regex:conv_wgrad_umma -c 4 -f -o gpurun_out/prof_wgrad python scripts/ncu_conv.py 0 > gpurun_out/ncu3.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out/ | grep ncu-rep
