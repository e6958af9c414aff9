#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cloud.py tests/test_gpu_icp.py -q > gpurun_out/k_pytest.log 2>&1; echo "exit $?" >> gpurun_out/k_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k_smoke.log 2>&1
timeout 300 python tools/k4_probe.py > gpurun_out/k_k4_probe.jsonl 2>&1
timeout 300 python tools/k4_probe.py --voxel 0.02 >> gpurun_out/k_k4_probe.jsonl 2>&1
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k_launches_k4_warm.csv \
  python tools/k4_probe.py --reps 2 --flush 0 > gpurun_out/k_ncu_launches_k4.log 2>&1
