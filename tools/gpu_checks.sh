#!/bin/bash
# Everything measured on a B200 box for profiles/: parity tests, smoke, bench lines, ncu launch lists and --set full
# captures.  Run from the repo root (e.g. `gpurun --timeout 2400 -- 'bash tools/gpu_checks.sh'`); writes gpurun_out/r02_*.
set -x
mkdir -p gpurun_out
O=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > ${O}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > ${O}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> ${O}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1
(while true; do nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv,noheader; sleep 0.5; done) > ${O}_clocks.csv &
CLK=$!
timeout 900 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err
kill $CLK
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference.json 2>&1
timeout 600 python tools/bench_kernels.py > ${O}_bench_kernels.jsonl 2> ${O}_bench_kernels.err
timeout 300 python tools/k4_probe.py > ${O}_k4_probe.jsonl 2>&1
timeout 300 python tools/k4_probe.py --voxel 0.02 >> ${O}_k4_probe.jsonl 2>&1
timeout 300 python tools/icp_probe.py > ${O}_icp_probe.jsonl 2>&1
timeout 300 python tools/icp_probe.py --host-loop >> ${O}_icp_probe.jsonl 2>&1
timeout 300 python tools/knn_probe.py > ${O}_knn_probe.jsonl 2>&1
for c in "bgr unit" "bgr packed8" "nv12 unit" "nv12 packed8" "bgr unit --r-max 0"; do set -- $c; timeout 300 python tools/k1_probe.py --color $1 --colors $2 $3 $4; done > ${O}_k1_probe.jsonl 2>&1
# ncu passes (never a bench value): launch lists (cold, serialised), warm K4 list, then one full capture per kernel family
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file ${O}_launches_bench_k1.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-rows --frames 4096 > ${O}_ncu_launches.log 2>&1
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_k4_warm.csv \
  python tools/k4_probe.py --reps 2 --flush 0 > ${O}_ncu_launches_k4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deproject -s 2 -c 1 -f -o ${O}_k1 \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-rows --frames 4096 > ${O}_ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deproject -s 4 -c 1 -f -o ${O}_k1_nv12 \
  python tools/k1_probe.py --color nv12 --colors packed8 --reps 2 > ${O}_ncu_k1_nv12.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_vox|k_transform" -c 12 -f -o ${O}_k34 \
  python tools/k4_probe.py --reps 1 > ${O}_ncu_k34.log 2>&1
timeout 600 ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_icp_warm.csv \
  python tools/icp_probe.py --reps 1 > ${O}_ncu_launches_icp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_nn_search|k_icp|k_knn_query" -c 12 -f -o ${O}_icp \
  python tools/icp_probe.py --reps 1 > ${O}_ncu_icp.log 2>&1
# gpurun brings back at most 64 MiB: condense the captures here, keep only the headline kernel's report itself
for r in k1 k1_nv12 k34 icp; do python tools/ncu_summary.py ${O}_$r.ncu-rep > ${O}_${r}_ncu_raw_subset.csv 2>> ${O}_ncu_summary.err; done
ncu -i ${O}_k1.ncu-rep --page source --csv 2>/dev/null | gzip > ${O}_k1_source.csv.gz
rm -f ${O}_k1_nv12.ncu-rep ${O}_k34.ncu-rep ${O}_icp.ncu-rep
du -sh gpurun_out
echo done
