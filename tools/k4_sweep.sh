#!/bin/bash
# A/B of the K4 tuning variants (tools/k1_variants.py k4_*): 5 mm and 20 mm grids on the benchmark's four-view cloud.
mkdir -p gpurun_out
for so in build/variants/librv_k4*.so; do
  n=$(basename $so .so)
  for v in 0.005 0.02; do
    echo -n "$n $v " ; RV_LIBRARY_PATH=$PWD/$so timeout 300 python tools/k4_probe.py --voxel $v 2>&1 | tail -1
  done
done | tee gpurun_out/k4_sweep.txt
