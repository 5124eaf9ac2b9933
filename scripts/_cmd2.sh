#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -3
SECONDS=0
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_full_n2.json 2> gpurun_out/bench_full_n2.err
echo "bench n2 wall: $SECONDS s"; tail -3 gpurun_out/bench_full_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'])
print('errors', d.get('extras_errors'))
print('per_rank', d.get('per_rank'))
print('exchange', json.dumps(d.get('exchange')))
print('strong', d.get('scaling_strong'))
PY
