#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2 3; do
timeout 60 python bench.py --steps 100 --warmup 5 --cpu-rows 2048 --no-extras 2>/dev/null | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print("merged: ms_per_step %.4f kernel_ms %.4f launches %s mismatches %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d.get("parity",{}).get("index_mismatches_vs_oracle")))'
GQ_B200_LIB=$PWD/scripts/_build/libgq_v0.so timeout 60 python bench.py --steps 100 --warmup 5 --cpu-rows 2048 --no-extras 2>/dev/null | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print("separate: ms_per_step %.4f kernel_ms %.4f launches %s mismatches %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d.get("parity",{}).get("index_mismatches_vs_oracle")))'
done
timeout 100 python tests/diag/graph_probe.py 2>&1 | tail -3
