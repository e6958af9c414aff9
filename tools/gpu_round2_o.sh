#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/icp_probe.py > gpurun_out/o_icp.jsonl 2>&1
timeout 300 python tools/icp_probe.py --host-loop >> gpurun_out/o_icp.jsonl 2>&1
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/o_icp_launches.csv \
  python tools/icp_probe.py --reps 1 > gpurun_out/o_ncu_icp.log 2>&1
