#!/bin/bash
for hold in 1 0 1 0 1 0 1 0 1 0 1 0 1 0 1 0; do
GQ_BENCH_E2E_HOLD=$hold timeout 60 python bench.py --steps 20 --warmup 5 --cpu-rows 256 --no-extras 2>/dev/null | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print("hold='$hold': ms_per_step %.4f e2e_ms %.4f" % (d["ms_per_step"], 65536/d["e2e"]["value"]*1e3))'
done
