#!/bin/bash
# On a B200 box (gpurun -- 'bash scripts/validate_gpu.sh'): GPU suite, the driver's bench command, the ncu launch list of
# the same command, smoke().  Writes gpurun_out/r02_*; scripts/collect_profiles.py r02 turns them into profiles/.
mkdir -p gpurun_out
SECONDS=0
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "tests wall: $SECONDS s"
SECONDS=0
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$? wall: $SECONDS s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])
print('parity', d.get('parity')); print('errors', d.get('extras_errors')); print('clocks', d.get('clocks'))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 240 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 30 --warmup 5 --no-extras --cpu-rows 256 > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/ncu_bench.log').read().strip().splitlines()[-1]); print('under ncu (not a bench value): ms_per_step', d['ms_per_step'])"
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
