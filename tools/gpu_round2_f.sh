#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cloud.py -q -k "voxel or fusion or smoke" > gpurun_out/f_pytest.log 2>&1; echo "exit $?" >> gpurun_out/f_pytest.log
bash tools/k4_sweep.sh > gpurun_out/f_k4_sweep.log 2>&1
for n in k1_publast k1_pubw0 k1_publast k1_pubw0; do
  RV_LIBRARY_PATH=$PWD/build/variants/librv_$n.so timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-rows --steps 10 --warmup 3 --frames 4096 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', round(d['value']), round(d['roofline']['frac'], 4))"
done > gpurun_out/f_k1_pub.txt 2>&1
echo done
