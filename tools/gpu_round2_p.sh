#!/bin/bash
# ICP loop after the fused move / warm-started search / register-resident solve: tests, probe, warm launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_neighbors.py -x -q -m gpu > gpurun_out/p_pytest.log 2>&1; echo "exit $?" >> gpurun_out/p_pytest.log
timeout 300 python tools/icp_probe.py > gpurun_out/p_icp.jsonl 2>&1
timeout 300 python tools/icp_probe.py --host-loop >> gpurun_out/p_icp.jsonl 2>&1
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_icp_launches.csv \
  python tools/icp_probe.py --reps 1 > gpurun_out/p_ncu_icp.log 2>&1
tail -3 gpurun_out/p_pytest.log; cat gpurun_out/p_icp.jsonl
