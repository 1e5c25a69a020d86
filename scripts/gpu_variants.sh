#!/bin/bash
# Steady-state throughput of kernel variants (zombie_b200/variants/lib*.so, selected with ZOMBIE_B200_LIB)
pts=${1:-8192}
true
for v in zombie_b200/variants/lib*.so; do echo "== $v"; ZOMBIE_B200_LIB=$PWD/$v ZB_REPS=1 timeout 120 python scripts/profile_case.py 8 $pts 64 2>&1 | tail -1; done
