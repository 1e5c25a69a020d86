#!/bin/bash
# Builds the A/B variants of the library measured on the GPU box (run here, on the CPU box; the .so files travel with the
# snapshot and are selected with ZOMBIE_B200_LIB). Round 2: occupancy of the two query kernels of the wavefront — smaller
# shared work lists (more CTAs per SM by shared memory) combined with a register cap (more CTAs per SM by registers).
set -e
cd "$(dirname "$0")/.."
python -m zombie_b200.build --variant occ5 -DZB_NODE_STACK=768 -DZB_PRIM_QUEUE=832 &
python -m zombie_b200.build --variant occ6 -DZB_NODE_STACK=512 -DZB_PRIM_QUEUE=704 -DZB_STAR_MIN_BLOCKS=6 -DZB_CLOSEST_MIN_BLOCKS=6 &
wait
python -m zombie_b200.build --variant occ7 -DZB_NODE_STACK=512 -DZB_PRIM_QUEUE=704 -DZB_STAR_MIN_BLOCKS=7 -DZB_CLOSEST_MIN_BLOCKS=7 &
python -m zombie_b200.build --variant occ8 -DZB_NODE_STACK=448 -DZB_PRIM_QUEUE=608 -DZB_STAR_MIN_BLOCKS=8 -DZB_CLOSEST_MIN_BLOCKS=8 &
wait
ls -la zombie_b200/variants/
