#!/bin/bash
tag=${1:-r2f}
out=gpurun_out/$tag
mkdir -p $out
A="gpurun_in 2.0 0.6832276 -0.09269869 0.530396 0.631202459 0.0"
echo "== trace"; timeout 60 tools/gpu_star_debug_trace $A 26 2>&1 | grep -v "^device\|^  prim" | tee $out/trace.log
for v in notwostage norange nojoint budget64 budget1 noplane noedgeflag; do echo "== $v"; timeout 60 tools/gpu_star_debug_$v $A 2>&1 | grep "different queries"; done 2>&1 | tee $out/bisect.log
for cfg in "ZB_WAVE_GROUPS=1 ZB_WAVE_STRIPE=1" "ZB_WAVE_GROUPS=1" "ZB_WAVE_GROUPS=4 ZB_WAVE_STRIPE=1" "ZB_WAVE_GROUPS=4" "ZB_SCHEDULER_PERSISTENT=1"; do echo "== $cfg"; env $cfg timeout 120 python scripts/wave_check.py 2>&1 | tail -2; done 2>&1 | tee $out/wave_check.log
