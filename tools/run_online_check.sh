set -x
mkdir -p gpurun_out
DX_ATTN_FWD64_2CTA=1 timeout 600 python -m pytest tests -m gpu -x -q -k "attention or attn" > gpurun_out/on_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/on_tests.log
tail -5 gpurun_out/on_tests.log
DX_ATTN_FWD64_2CTA=1 timeout 300 python tools/attn_time.py 2>&1 | grep fwd > gpurun_out/on_attn2.log
DX_ATTN_FWD64_2CTA=0 timeout 300 python tools/attn_time.py 2>&1 | grep fwd > gpurun_out/on_attn1.log
cat gpurun_out/on_attn2.log gpurun_out/on_attn1.log
DX_AB_ONLY=all_on,fwd64_2cta timeout 900 python tools/ab_step.py > gpurun_out/on_ab.log 2>&1
tail -3 gpurun_out/on_ab.log
