#!/bin/bash
# Same-box A/B of two builds of libclover_b200.so (power-capped clocks differ between boxes by +-3 %):
#   bash tools/ab_bench.sh build_ab/libclover_b200_base.so [extra bench flags]
# Writes gpurun_out/ab_{base,new}.json and gpurun_out/ab_{base,new}_family_times.txt, then prints the comparison.
BASE=$1; shift
mkdir -p gpurun_out
export CLOVER_B200_PROFILE_SHAPES=1
for round in 1 2; do
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --lib "$BASE" "$@" > gpurun_out/ab_base_$round.json 2> gpurun_out/ab_base.err
  mv gpurun_out/family_times.txt gpurun_out/ab_base_family_times.txt
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline "$@" > gpurun_out/ab_new_$round.json 2> gpurun_out/ab_new.err
  mv gpurun_out/family_times.txt gpurun_out/ab_new_family_times.txt
done
python tools/ab_compare.py
