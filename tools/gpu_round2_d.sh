#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/d_pytest.log 2>&1; echo "exit $?" >> gpurun_out/d_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/d_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "exit $?" >> gpurun_out/d_bench.err
timeout 600 python tools/bench_kernels.py > gpurun_out/d_bench_kernels.jsonl 2> gpurun_out/d_bench_kernels.err
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_vox" -c 12 -f -o gpurun_out/d_k4 \
  python tools/bench_kernels.py --reps 1 > gpurun_out/d_ncu_k4.log 2>&1
echo done
