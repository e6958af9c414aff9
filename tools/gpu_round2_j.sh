#!/bin/bash
mkdir -p gpurun_out
for cfg in "2048 2" "4096 1" "3072 1" "1024 2"; do set -- $cfg
  timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-rows --steps 10 --warmup 3 --frames 8192 --chunk $1 --ring-slots $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $1 slots $2', round(d['value']), round(d['roofline']['frac'], 4), d['roofline']['launch_ms'])"
done > gpurun_out/j_chunk.txt 2>&1
