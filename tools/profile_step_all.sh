#!/bin/bash
# Every launch of ONE eager training step (bench workload) under ncu with the sections that carry DRAM bytes, tensor-pipe activity,
# issue activity and warp-state statistics (no source import: ~420 launches).  Output: gpurun_out/r2_step_all.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis \
    --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis \
    --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active \
    -f -o gpurun_out/r2_step_all python tools/step_launches.py > gpurun_out/r2_step_all.log 2>&1
ls -la gpurun_out/r2_step_all.ncu-rep
tail -3 gpurun_out/r2_step_all.log
