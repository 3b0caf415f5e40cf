cd /root/repo
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
( time run --steps 20 --warmup 5 > gpurun_out/r02_bench_two_stage_${N}gpu.json ) 2> gpurun_out/b${N}.err; tail -3 gpurun_out/b${N}.err
( time run --steps 5 --workload single_stage_w6 --no-strong > gpurun_out/r02_bench_single_stage_${N}gpu.json ) 2> gpurun_out/b${N}s.err; tail -3 gpurun_out/b${N}s.err
( time run --workload pairs_c5 --steps 8 > gpurun_out/r02_bench_pairs_${N}gpu.json ) 2> gpurun_out/b${N}p.err; tail -3 gpurun_out/b${N}p.err
( time run --workload resample_c4 > gpurun_out/r02_bench_resample_${N}gpu.json ) 2> gpurun_out/b${N}r.err; tail -3 gpurun_out/b${N}r.err
python - <<P
import json
N=$N
d=json.load(open(f'gpurun_out/r02_bench_two_stage_{N}gpu.json')); print('two-stage', d['value'], d['ms_per_step'], d['parity'], json.dumps(d['strong']), d['e2e'])
d=json.load(open(f'gpurun_out/r02_bench_single_stage_{N}gpu.json')); print('single', d['value'], d['ms_per_step'], d['parity'])
d=json.load(open(f'gpurun_out/r02_bench_pairs_{N}gpu.json')); print('pairs', d['value'], d['pairs_per_s'], d['e2e'])
d=json.load(open(f'gpurun_out/r02_bench_resample_{N}gpu.json')); print('resample', d['single_stage'], d['two_stage'])
P
