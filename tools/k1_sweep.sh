#!/bin/bash
# A/B of the K1 tuning variants built by tools/k1_variants.py: headline workload, device-resident arm only.
mkdir -p gpurun_out
for so in build/variants/librv_*.so; do
  n=$(basename $so .so)
  RV_LIBRARY_PATH=$PWD/$so timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-rows --steps 8 --warmup 3 --frames 4096 \
    > gpurun_out/sweep_$n.json 2> gpurun_out/sweep_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"]), round(d["roofline"]["frac"], 4), round(d["roofline"]["launch_ms"], 4), "kept points/step", d["valid_points_per_step_rank0"])
except Exception as e:
    print("$n", "failed", e)
PY
done | tee gpurun_out/sweep_summary.txt
