#!/bin/bash
mkdir -p gpurun_out
bash tools/icp_sweep.sh > /dev/null 2>&1
RV_LIBRARY_PATH=$PWD/build/variants/librv_icp_c30_timing.so timeout 300 python tools/icp_probe.py --reps 1 --dump-state > gpurun_out/r_timing.txt 2>&1
cat gpurun_out/icp_sweep.txt gpurun_out/r_timing.txt
