#!/bin/bash
# Builds the A/B variants of the library measured by scripts/gpu_variants.sh (run here, on the CPU box; the .so files
# travel to the GPU box with the snapshot). Variants of round 1's last change (code shape) and of the phase-1 budget.
set -e
cd "$(dirname "$0")/.."
python -m zombie_b200.build --variant compact                                           # the default build
python -m zombie_b200.build --variant unrolled_inlined -DZB_UNROLL_CHILD_LOOPS -DZB_INLINE_PRIM_TESTS   # old code shape
python -m zombie_b200.build --variant unrolled -DZB_UNROLL_CHILD_LOOPS
python -m zombie_b200.build --variant unroll2 -DZB_CHILD_UNROLL=2
python -m zombie_b200.build --variant unroll4 -DZB_CHILD_UNROLL=4
python -m zombie_b200.build --variant inlined -DZB_INLINE_PRIM_TESTS
python -m zombie_b200.build --variant no_plane_bound -DZB_ROBIN_PLANE_BOUND=0 -DZB_STAR_TWO_STAGE=0
python -m zombie_b200.build --variant one_stage -DZB_STAR_TWO_STAGE=0
python -m zombie_b200.build --variant no_closest_filter -DZB_CLOSEST_PLANE_FILTER=0
python -m zombie_b200.build --variant no_edge_flag -DZB_NO_EDGE_FLAG=0
python -m zombie_b200.build --variant budget4 -DZB_PHASE1_BUDGET=4
python -m zombie_b200.build --variant budget16 -DZB_PHASE1_BUDGET=16
ls -la zombie_b200/variants/
