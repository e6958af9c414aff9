#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deproject.py -q -x > gpurun_out/c_pytest_k1.log 2>&1; echo "exit $?" >> gpurun_out/c_pytest_k1.log
bash tools/k1_sweep.sh > gpurun_out/c_sweep.log 2>&1
for c in "bgr unit" "bgr packed8" "nv12 unit" "nv12 packed8"; do set -- $c; timeout 300 python tools/k1_probe.py --color $1 --colors $2; done > gpurun_out/c_probe.jsonl 2> gpurun_out/c_probe.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deproject -s 4 -c 1 -f -o gpurun_out/c_k1_nv12 \
  python tools/k1_probe.py --color nv12 --colors packed8 --reps 2 > gpurun_out/c_ncu_nv12.log 2>&1
echo done
