STEPS=200 bash scripts/run_variants.sh
STEPS=200 bash scripts/run_variants.sh
GQ_B200_LIB=$PWD/scripts/_build/libgq_v1.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quantize or module_eval or gq2 or fused or search_modes or graphed" 2>&1 | tail -3
