#!/bin/bash
# ncu --set full captures of the step's heaviest kernels (one launch each, after one warm launch)
M='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|gpu__dram_throughput.avg.pct|sm__pipe_tensor_cycles_active|sm__warps_active.avg.pct|launch__registers|lts__t_sector_hit_rate.pct|smsp__average_warps_issue_stalled|smsp__warp_issue_stalled.*_per_warp_active|sm__throughput.avg.pct|l1tex__m_xbar2l1tex_read_bytes|lts__throughput.avg.pct|launch__grid_size|launch__occupancy_limit|sm__inst_executed_pipe_lsu|smsp__cycles_active.avg|sm__cycles_elapsed.max'
for k in conv_fwd_halo epilogue_bwd_vec blur_tma; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 -c 3 -f -o gpurun_out/r01_$k python scripts/ncu_step_kernels.py > gpurun_out/ncu_$k.log 2>&1; echo "$k rc=$?"
  ncu -i gpurun_out/r01_$k.ncu-rep --page raw --csv > gpurun_out/r01_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r01_$k.ncu-rep --page source --csv > gpurun_out/r01_${k}_source.csv 2>/dev/null
done
ls -la gpurun_out | tail -20
