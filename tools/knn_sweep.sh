#!/bin/bash
# A/B of the k-nearest-neighbour query variants (tools/k1_variants.py knn_*) on tools/knn_probe.py
mkdir -p gpurun_out
for so in build/variants/librv_knn_*.so; do
  n=$(basename $so .so)
  echo -n "$n " ; RV_LIBRARY_PATH=$PWD/$so timeout 300 python tools/knn_probe.py 2>&1 | tail -1
done | tee gpurun_out/knn_sweep.txt
