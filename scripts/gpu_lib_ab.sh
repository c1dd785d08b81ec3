#!/bin/bash
# steady-state A/B of library builds: gpu_lib_ab.sh "<nq> <qtok> <pages>" libA libB ...
CASE=$1; shift
mkdir -p gpurun_out
for round in 1 2; do
  for lib in "$@"; do
    echo "== round $round lib $lib case $CASE"
    LIS_LIB=multi-modal_colpali_b200/_lib/$lib timeout 300 python scripts/gpu_pair_ab2.py $CASE 2>&1 | grep '"mode": "pair"' | tail -2
  done
done
