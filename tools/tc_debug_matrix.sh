# DX_TC_DEBUG masks over the isolated GEMM timings (tools/tc_time_passes.py): which part of the kernel the time goes to
for m in ${DX_MASKS:-0 8 3 11 79 335}; do
  echo "== DX_TC_DEBUG=$m"
  DX_TC_DEBUG=$m timeout 120 python tools/tc_time_passes.py 2>&1 | sed -e 's/ | p2:.*//' | cut -c1-90
done
