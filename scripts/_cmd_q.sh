#!/bin/bash
timeout 200 python tests/diag/slow_batch_probe.py 8 2>&1 | tail -1
timeout 100 python tests/diag/stress_cap_probe.py 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
timeout 60 python bench.py --steps 20 --warmup 5 --cpu-rows 2048 --no-extras 2>/dev/null | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print("main: ms_per_step %.4f kernel_ms %.4f mismatches %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity",{}).get("index_mismatches_vs_oracle")))'
done
