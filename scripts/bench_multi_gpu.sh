#!/bin/bash
# The 2-rank GPU tests, then the full bench.py --gpus N line under torchrun (gpurun --gpus N -- 'bash scripts/bench_multi_gpu.sh N').
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8 | wc -l
[ "${SKIP_TESTS:-0}" = 1 ] || timeout 300 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -3
SECONDS=0
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_full_n$N.json 2> gpurun_out/bench_full_n$N.err
echo "bench n$N rc=$? wall: $SECONDS s"; grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/bench_full_n$N.err | tail -3 | cut -c1-300
python - $N <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_full_n%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])
print('errors', d.get('extras_errors'))
for k,v in d.get('exchange',{}).items():
    if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','collective_us','mode','equals_nccl_all_gather','equals_nccl_fold')})
    else: print(k, v)
print('strong', {k:d['scaling_strong'][k] for k in ('ms_per_step','value')})
PY
