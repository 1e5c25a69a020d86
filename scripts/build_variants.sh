#!/bin/bash
# Builds the A/B variants of the library measured on the GPU box (run here, on the CPU box; the .so files travel with the
# snapshot and are selected with ZOMBIE_B200_LIB). Round 2: occupancy of the two query kernels of the wavefront — the size
# of the shared work lists (CTAs per SM by shared memory) against a register cap (CTAs per SM by registers). The default
# build is lists 512/704 with six CTAs per SM; results in profiles/README.md.
set -e
cd "$(dirname "$0")/.."
python -m zombie_b200.build --variant lists1024 -DZB_NODE_STACK=1024 -DZB_PRIM_QUEUE=1120 -DZB_STAR_MIN_BLOCKS=1 -DZB_CLOSEST_MIN_BLOCKS=1 &
python -m zombie_b200.build --variant occ5 -DZB_NODE_STACK=768 -DZB_PRIM_QUEUE=832 -DZB_STAR_MIN_BLOCKS=1 -DZB_CLOSEST_MIN_BLOCKS=1 &
wait
python -m zombie_b200.build --variant occ7 -DZB_STAR_MIN_BLOCKS=7 -DZB_CLOSEST_MIN_BLOCKS=7 &
python -m zombie_b200.build --variant occ8 -DZB_NODE_STACK=448 -DZB_PRIM_QUEUE=608 -DZB_STAR_MIN_BLOCKS=8 -DZB_CLOSEST_MIN_BLOCKS=8 &
wait
ls -la zombie_b200/variants/
