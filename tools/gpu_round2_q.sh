#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_register.py -x -q -m gpu > gpurun_out/q_pytest.log 2>&1; echo "exit $?" >> gpurun_out/q_pytest.log
bash tools/icp_sweep.sh > /dev/null 2>&1
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_icp_launches.csv \
  python tools/icp_probe.py --reps 1 > gpurun_out/q_ncu_icp.log 2>&1
tail -3 gpurun_out/q_pytest.log; cat gpurun_out/icp_sweep.txt
