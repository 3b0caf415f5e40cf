cd /root/repo
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json ) 2> gpurun_out/bench_n1.err
tail -5 gpurun_out/bench_n1.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','parity')})
print('roofline', {k:d['roofline'][k] for k in ('frac','frac_whole_step','frac_whole_step_sustained','kernel_ms','kernels_ms_per_step')})
print('e2e', d['e2e']); print('cpu', d['cpu_baseline'])
print('config', d['config'].get('ms_per_step_sustained'))
print(json.dumps(d['extra_workloads'], indent=1)[:6000])
P
