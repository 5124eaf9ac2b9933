#!/bin/bash
# diagnostic: run bench.py against each scripts/_build/libgq_v*.so and print the search-kernel time
cd "$(dirname "$0")/.."
for f in scripts/_build/libgq_v*.so; do
  out=$(GQ_B200_LIB=$PWD/$f timeout 60 python bench.py --steps ${STEPS:-60} --warmup 5 --cpu-rows 2048 --no-extras 2>/dev/null | tail -1)
  echo "$f $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("ms_per_step %.4f kernel_ms %.4f frac %.3f mismatches %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d.get("parity",{}).get("index_mismatches_vs_oracle")))
except Exception as e: print("FAILED", e)')"
done
