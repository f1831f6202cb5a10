#!/bin/bash
# Round-2 profile capture (run under gpurun on ONE GPU): launch list of one eager training step, aggregated per kernel.
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv python tools/step_launches.py > gpurun_out/r2_sl.log 2>&1
python tools/agg_launches.py gpurun_out/r2_launches.csv 60 > gpurun_out/r2_agg.txt 2>&1
cat gpurun_out/r2_agg.txt
