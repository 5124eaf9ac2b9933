#!/bin/bash
timeout 100 python tests/diag/graph_probe.py 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_segment or fused_latent or module_eval or stress" 2>&1 | tail -2
