#!/bin/bash
# Round profile capture (run under gpurun on ONE GPU): launch list of one eager training step + ncu --set full captures of
# the dominant GEMM and of the tcgen05 attention forward / backward.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/step_launches.py > gpurun_out/sl.log 2>&1
python tools/agg_launches.py gpurun_out/launches.csv 40 > gpurun_out/agg.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_conv1 \
    python tools/dominant_kernel.py > gpurun_out/prof_conv1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc_kernel -s 3 -c 2 -f -o gpurun_out/prof_attn \
    python tools/attn_time.py > gpurun_out/prof_attn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 6 -c 2 -f -o gpurun_out/prof_attn_bwd \
    python tools/attn_time.py > gpurun_out/prof_attn_bwd.log 2>&1
ls -la gpurun_out
cat gpurun_out/agg.txt
