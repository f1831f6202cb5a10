#!/bin/bash
# gpurun_out/r2_prof_*.ncu-rep (tools/profile_round2_full.sh) -> profiles/r2_*.json ; r2_step_all.ncu-rep -> profiles/r2_step_kernels.{json,md}
set -e
python tools/ncu_step_summary.py gpurun_out/r2_step_all.ncu-rep profiles/r2_step_kernels.json profiles/r2_step_kernels.md > /dev/null
python tools/ncu_summary.py gpurun_out/r2_prof_conv1.ncu-rep profiles/r2_dominant_kernel.json "gemm_tc_kernel<BF16X3, HALO, STD>, FFT conv1 (32x1000 rows, 128 -> 1024, k=3, planes-only output), as bench.py times it"
for k in wgrad inproj conv2ln outprojln attn_fwd16 attn_fwd64 attn_bwd16 attn_bwd64 gauss_fwd ln_bwd; do
  if [ -f gpurun_out/r2_prof_$k.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/r2_prof_$k.ncu-rep profiles/r2_$k.json "one launch inside the bench training step (tools/step_launches.py), kernel filter: $k"
  fi
done
