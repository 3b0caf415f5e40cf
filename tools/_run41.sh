set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipeline or peer_exchange or two_stage" 2>&1 | tail -5
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_two_stage_v2.json 2> gpurun_out/bench41.err; tail -3 gpurun_out/bench41.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_two_stage_v2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print(d['config']['pipelined']); print(d['config']['single_stack'])
print({k:v for k,v in d['roofline'].items() if k!='kernels_ms_per_step'})
print(d['e2e']); print(d['cpu_baseline']); print(d['parity'])
for k,v in (d['extra_workloads'] or {}).items(): print(k, json.dumps(v)[:900])
PY
