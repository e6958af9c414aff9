#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/knn_probe.py > gpurun_out/w_knn.jsonl 2>&1
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/w_knn_launches.csv \
  python tools/knn_probe.py --reps 1 > gpurun_out/w_ncu_knn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_query -c 4 -o gpurun_out/w_knn_query \
  python tools/knn_probe.py --reps 1 > gpurun_out/w_ncu_full.log 2>&1
cat gpurun_out/w_knn.jsonl
