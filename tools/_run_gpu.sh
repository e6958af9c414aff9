set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_voxel" -c 6 -f -o gpurun_out/k4_v6 python tools/bench_kernels.py --reps 1 > gpurun_out/ncu_k4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_all.csv python tools/bench_kernels.py --reps 1 > gpurun_out/ncu_launches3.log 2>&1
echo done
