#!/bin/bash
# A/B of the ICP search variants (tools/k1_variants.py icp_*) on tools/icp_probe.py's registration
mkdir -p gpurun_out
for so in build/variants/librv_icp_*.so; do
  n=$(basename $so .so)
  echo -n "$n " ; RV_LIBRARY_PATH=$PWD/$so timeout 300 python tools/icp_probe.py 2>&1 | tail -1
done | tee gpurun_out/icp_sweep.txt
