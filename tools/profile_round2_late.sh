#!/bin/bash
# Late round-2 evidence (one GPU under gpurun): the step launch list again after the one-pass / three-tap weight gradients and the one-pass
# attention softmax, plus --set full captures of the new kernels.  Reports stay on the box; summaries come back in gpurun_out/.
P=/tmp/dxprof
mkdir -p gpurun_out $P
timeout 900 ncu --profile-from-start off --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis \
    --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis \
    --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active \
    -f -o $P/r2b_step_all python tools/step_launches.py > gpurun_out/r2b_step_all.log 2>&1
python tools/ncu_step_summary.py $P/r2b_step_all.ncu-rep gpurun_out/r2b_step_kernels.json gpurun_out/r2b_step_kernels.md > /dev/null 2>gpurun_out/r2b_step_summary.err
full() { # name regex skip
  timeout 400 ncu --kernel-name-base demangled --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f \
      -o $P/r2b_prof_$1 python tools/step_launches.py > gpurun_out/r2b_prof_$1.log 2>&1
  python tools/ncu_summary.py $P/r2b_prof_$1.ncu-rep gpurun_out/r2b_$1.json "one launch inside the bench training step (tools/step_launches.py); kernel filter: $2, skip $3"
}
full wgrad3 'wgrad_halo3_kernel' 1
full attn_fwd64 'attn_fwd_tc_kernel<\(int\)64' 4
ls -la gpurun_out | tail -12
