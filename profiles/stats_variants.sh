#!/bin/sh
# Build-flag sweep of the tile-stat kernel (run on the GPU box from the repo root): occupancy vs registers vs loads in flight.
set -e
for v in "-DFAST_MIN_BLOCKS=4 -DFAST_UNROLL_N=4" "-DFAST_MIN_BLOCKS=5 -DFAST_UNROLL_N=4" "-DFAST_MIN_BLOCKS=6 -DFAST_UNROLL_N=4" "-DFAST_MIN_BLOCKS=5 -DFAST_UNROLL_N=2" "-DFAST_MIN_BLOCKS=6 -DFAST_UNROLL_N=2" "-DFAST_MIN_BLOCKS=4 -DFAST_UNROLL_N=8" "-DFAST_MIN_BLOCKS=3 -DFAST_UNROLL_N=8"; do
  QA_NVCC_EXTRA="$v" sh quantization_analysis_b200/csrc/build.sh > /dev/null 2>&1
  printf "%s : " "$v"
  python profiles/step_parts.py 2>/dev/null | grep "^stats  "
done
sh quantization_analysis_b200/csrc/build.sh > /dev/null 2>&1
