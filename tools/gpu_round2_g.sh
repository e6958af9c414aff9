#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "exit $?" >> gpurun_out/g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1
bash tools/k4_sweep.sh > gpurun_out/g_k4_sweep.log 2>&1
timeout 900 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "exit $?" >> gpurun_out/g_bench.err
echo done
