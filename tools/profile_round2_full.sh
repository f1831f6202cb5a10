#!/bin/bash
# Round-2 evidence (run under gpurun on ONE GPU).  The .ncu-rep files stay on the box (/tmp/dxprof); only the JSON / markdown summaries
# (tools/ncu_step_summary.py, tools/ncu_summary.py run there) come back through gpurun_out/ (64 MiB limit).
#   1. every launch of one eager training step with DRAM bytes / tensor-pipe / issue / warp-state sections  -> r2_step_kernels.{json,md}
#   2. ncu --set full --import-source on, ONE launch each, of: the dominant conv-GEMM (as bench.py times it), the conv1 weight-gradient GEMM,
#      the in-projection GEMM with the head-plane epilogue, conv2 / out-projection + LayerNorm epilogue, the four attention kernels,
#      Gaussian upsampling forward, LayerNorm backward                                                         -> r2_<name>.json
P=/tmp/dxprof
mkdir -p gpurun_out $P
timeout 900 ncu --profile-from-start off --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis \
    --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis \
    --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active \
    -f -o $P/r2_step_all python tools/step_launches.py > gpurun_out/r2_step_all.log 2>&1
python tools/ncu_step_summary.py $P/r2_step_all.ncu-rep gpurun_out/r2_step_kernels.json gpurun_out/r2_step_kernels.md > /dev/null 2>gpurun_out/r2_step_summary.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o $P/r2_prof_conv1 \
    python tools/dominant_kernel.py > gpurun_out/r2_prof_conv1.log 2>&1
python tools/ncu_summary.py $P/r2_prof_conv1.ncu-rep gpurun_out/r2_dominant_kernel.json "gemm_tc_kernel<BF16X3, HALO, STD>, FFT conv1 (32x1000 rows, 128 -> 1024, k=3, planes-only output), as bench.py times it"
full() { # name regex skip
  timeout 400 ncu --kernel-name-base demangled --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f \
      -o $P/r2_prof_$1 python tools/step_launches.py > gpurun_out/r2_prof_$1.log 2>&1
  python tools/ncu_summary.py $P/r2_prof_$1.ncu-rep gpurun_out/r2_$1.json "one launch inside the bench training step (tools/step_launches.py); kernel filter: $2, skip $3"
}
full wgrad 'gemm_tc_kernel<\(int\)1, \(int\)1' 1
full inproj 'gemm_tc_kernel<\(int\)1, \(int\)0, \(int\)0' 0
full conv2ln 'gemm_tc_kernel<\(int\)1, \(int\)2, \(int\)1' 0
full outprojln 'gemm_tc_kernel<\(int\)1, \(int\)0, \(int\)1' 0
full attn_fwd16 'attn_fwd_tc_kernel<\(int\)16' 0
full attn_fwd64 'attn_fwd_tc_kernel<\(int\)64' 4
full attn_bwd16 'attn_bwd_tc_pipe_kernel' 0
full attn_bwd64 'attn_bwd_tc_pipe64_kernel' 0
full gauss_fwd 'gauss_upsample_fwd_kernel' 0
full ln_bwd 'ln_bwd_fused_kernel<\(int\)4' 0
# warp-stall breakdown per source line for the two attention backward kernels: top SASS lines by samples
for k in attn_bwd16 attn_bwd64 attn_fwd16; do
  ncu -i $P/r2_prof_$k.ncu-rep --page source --csv 2>/dev/null | python tools/ncu_top_stalls.py > gpurun_out/r2_${k}_stalls.txt 2>/dev/null
done
du -sh gpurun_out; ls gpurun_out | head -50
