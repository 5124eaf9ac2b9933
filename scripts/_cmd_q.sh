#!/bin/bash
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_peer.py -m gpu -x -q -k "not times_out" 2>&1 | tail -6
echo "memcheck rc=$?"
timeout 280 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_peer.py -m gpu -x -q -k "store_kernel or kl_fold" 2>&1 | tail -6
echo "racecheck rc=$?"
