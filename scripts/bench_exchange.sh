#!/bin/bash
# bench.py --gpus N --no-extras under torchrun (gpurun --gpus N -- 'bash scripts/bench_exchange.sh N'): prints the
# headline and the exchange records (NCCL and peer-memory), writes gpurun_out/bench_peer_nN.json.
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/bench_peer_n$N.json 2> gpurun_out/bench_peer_n$N.err
echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/bench_peer_n$N.err | tail -5 | cut -c1-300
python - $N <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_peer_n%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')})
for k,v in d.get('exchange',{}).items():
    if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('note','unit')})
    else: print(k, v)
PY
