#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -8
