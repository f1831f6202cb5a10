#!/bin/bash
# Round-2 evidence (run under gpurun on ONE GPU):
#   1. every launch of one eager training step with DRAM bytes / tensor-pipe / issue / warp-state sections  -> r2_step_all.ncu-rep
#   2. ncu --set full --import-source on, ONE launch each, of: the dominant conv-GEMM (as bench.py times it), the conv1 weight-gradient GEMM,
#      the in-projection GEMM with the head-plane epilogue, conv2 + LayerNorm epilogue, the four attention kernels
mkdir -p gpurun_out
bash tools/profile_step_all.sh
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_conv1 \
    python tools/dominant_kernel.py > gpurun_out/r2_prof_conv1.log 2>&1
# inside the training step: kernel-id filters pick one launch of each (skip counts land on frame-side, T = 1000 layers)
full() { # name regex skip
  timeout 400 ncu --kernel-name-base demangled --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/r2_prof_$1 \
      python tools/step_launches.py > gpurun_out/r2_prof_$1.log 2>&1
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
ls -la gpurun_out/r2_prof_*.ncu-rep
