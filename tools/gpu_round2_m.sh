#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deproject.py -q -x > gpurun_out/m_pytest_k1.log 2>&1; echo "exit $?" >> gpurun_out/m_pytest_k1.log
for n in k1_pre1 k1_pre0 k1_pre1 k1_pre0; do
  RV_LIBRARY_PATH=$PWD/build/variants/librv_$n.so timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-rows --steps 10 --warmup 3 --frames 4096 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', round(d['value']), round(d['roofline']['frac'], 4), d['valid_points_per_step_rank0'])"
done > gpurun_out/m_k1_pre.txt 2>&1
for n in k1_pre1 k1_pre0; do for c in "bgr unit --r-max 0" "nv12 packed8" "bgr unit --r-max 1.5"; do set -- $c; echo -n "$n "; RV_LIBRARY_PATH=$PWD/build/variants/librv_$n.so timeout 300 python tools/k1_probe.py --color $1 --colors $2 $3 $4 | tail -1; done; done > gpurun_out/m_k1_probe.txt 2>&1
