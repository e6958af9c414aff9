#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_icp.py tests/test_gpu_register.py tests/test_gpu_cloud.py -x -q -m gpu > gpurun_out/u_pytest.log 2>&1; echo "exit $?" >> gpurun_out/u_pytest.log
timeout 300 python tools/icp_probe.py > gpurun_out/u_icp.jsonl 2>&1
timeout 300 python tools/icp_probe.py --host-loop >> gpurun_out/u_icp.jsonl 2>&1
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/u_icp_launches.csv \
  python tools/icp_probe.py --reps 1 > gpurun_out/u_ncu_icp.log 2>&1
tail -3 gpurun_out/u_pytest.log; cat gpurun_out/u_icp.jsonl
