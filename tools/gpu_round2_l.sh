#!/bin/bash
mkdir -p gpurun_out
for r in 0.01 0.7 1.0 1.5; do timeout 300 python tools/k1_probe.py --color bgr --colors unit --r-max $r | tail -1; done > gpurun_out/l_k1_rmax.jsonl 2>&1
timeout 300 python tools/k1_probe.py --color nv12 --colors packed8 --r-max 0.01 | tail -1 >> gpurun_out/l_k1_rmax.jsonl 2>&1
