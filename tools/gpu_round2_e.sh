#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_cloud.py tests/test_gpu_icp.py -q > gpurun_out/e_pytest.log 2>&1; echo "exit $?" >> gpurun_out/e_pytest.log
timeout 300 python tools/k4_probe.py > gpurun_out/e_k4_probe.jsonl 2>&1
timeout 300 python tools/k4_probe.py --voxel 0.02 >> gpurun_out/e_k4_probe.jsonl 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/e_k4_launches.csv \
  python tools/k4_probe.py --reps 2 --flush 0 > gpurun_out/e_ncu_k4.log 2>&1
echo done
