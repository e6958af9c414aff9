#!/bin/bash
mkdir -p gpurun_out
bash tools/icp_sweep.sh > /dev/null 2>&1
bash tools/icp_sweep.sh > /dev/null 2>&1
cat gpurun_out/icp_sweep.txt
