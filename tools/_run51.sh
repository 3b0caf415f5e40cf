cd /root/repo
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
( time run --steps 20 --warmup 5 > gpurun_out/r02_bench_two_stage_${N}gpu_v2.json ) 2> gpurun_out/b${N}.err; tail -5 gpurun_out/b${N}.err
python - <<P
import json
N=$N
d=json.loads(open(f'gpurun_out/r02_bench_two_stage_{N}gpu_v2.json').read().strip().splitlines()[-1]); print('two-stage', d['value'], d['ms_per_step'], d['parity']); print(d['config']['pipelined']); print(d['config']['single_stack']); print(json.dumps(d['strong'])[:600]); print(d['e2e'])
P
