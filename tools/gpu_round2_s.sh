#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_icp.py tests/test_gpu_register.py tests/test_gpu_cloud.py -x -q -m gpu > gpurun_out/s_pytest.log 2>&1; echo "exit $?" >> gpurun_out/s_pytest.log
timeout 300 python tools/icp_probe.py > gpurun_out/s_icp.jsonl 2>&1
timeout 300 python tools/icp_probe.py --host-loop >> gpurun_out/s_icp.jsonl 2>&1
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s_icp_launches.csv \
  python tools/icp_probe.py --reps 1 > gpurun_out/s_ncu_icp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nn_search -c 3 -o gpurun_out/s_nn_search \
  python tools/icp_probe.py --reps 1 > gpurun_out/s_ncu_full.log 2>&1
tail -3 gpurun_out/s_pytest.log; cat gpurun_out/s_icp.jsonl
