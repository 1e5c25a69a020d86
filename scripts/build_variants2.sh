#!/bin/bash
# second batch of occupancy variants around the best of the first (occ6: lists 512/704, six CTAs per SM)
set -e
cd "$(dirname "$0")/.."
L="-DZB_NODE_STACK=512 -DZB_PRIM_QUEUE=704"
python -m zombie_b200.build --variant occ6 $L -DZB_STAR_MIN_BLOCKS=6 -DZB_CLOSEST_MIN_BLOCKS=6 &
python -m zombie_b200.build --variant occ6c7 $L -DZB_STAR_MIN_BLOCKS=6 -DZB_CLOSEST_MIN_BLOCKS=7 &
wait
python -m zombie_b200.build --variant occ6b4 $L -DZB_STAR_MIN_BLOCKS=6 -DZB_CLOSEST_MIN_BLOCKS=6 -DZB_PHASE1_BUDGET=4 &
python -m zombie_b200.build --variant occ6b16 $L -DZB_STAR_MIN_BLOCKS=6 -DZB_CLOSEST_MIN_BLOCKS=6 -DZB_PHASE1_BUDGET=16 &
wait
ls -la zombie_b200/variants/
