#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "exit $?" >> gpurun_out/b_pytest.log
bash tools/k1_sweep.sh > gpurun_out/b_sweep.log 2>&1
echo done
