#!/bin/bash
# round 2, first GPU pass: parity of the reworked K1 (mbarrier totals, whole-run rejection, NV12 in, packed colours out),
# then the bench line and the K1 captures.  Writes gpurun_out/a_*.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_deproject.py -x -q > gpurun_out/a_pytest_k1.log 2>&1; echo "exit $?" >> gpurun_out/a_pytest_k1.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_deproject.py > gpurun_out/a_pytest_rest.log 2>&1; echo "exit $?" >> gpurun_out/a_pytest_rest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "exit $?" >> gpurun_out/a_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/a_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-rows --frames 4096 > gpurun_out/a_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deproject -s 2 -c 1 -f -o gpurun_out/a_k1 \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-rows --frames 4096 > gpurun_out/a_ncu_k1.log 2>&1
echo done
