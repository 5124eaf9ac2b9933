#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
SECONDS=0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_full_n$N.json 2> gpurun_out/bench_full_n$N.err
echo "bench n$N rc=$? wall: $SECONDS s"; grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/bench_full_n$N.err | tail -5 | cut -c1-300
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f'gpurun_out/bench_full_n{N}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'frac', d['roofline']['frac'], 'kms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['config'].get('host_numa_binding'))
print('errors', d.get('extras_errors'))
print('per_rank', json.dumps(d.get('per_rank')))
print('exchange', json.dumps(d.get('exchange')))
print('strong', json.dumps(d.get('scaling_strong')))
PY
SECONDS=0
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
echo "ref arm wall: $SECONDS s"
