#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -8
bash scripts/_cmd_peer.sh 2
