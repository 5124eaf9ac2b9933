#!/bin/bash
export GQ_B200_LIB=$PWD/scripts/_build/libgq_v0.so
for i in 1 2; do
timeout 60 python bench.py --steps 100 --warmup 5 --cpu-rows 2048 --no-extras 2>/dev/null | tail -1 | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("TS: ms_per_step %.4f kernel_ms %.4f frac %.3f mismatches %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d.get("parity",{}).get("index_mismatches_vs_oracle")))
except Exception as e: print("FAILED", e)'
done
unset GQ_B200_LIB
GQ_PROF_TS=1 timeout 60 python scripts/tc_prof.py 2>&1 | tail -80
