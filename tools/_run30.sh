cd /root/repo
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( time run 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_two_stage_2gpu.json ) 2> gpurun_out/b2.err; tail -3 gpurun_out/b2.err
( time run 2 --steps 5 --workload single_stage_w6 --no-strong > gpurun_out/r02_bench_single_stage_2gpu.json ) 2> gpurun_out/b2s.err; tail -3 gpurun_out/b2s.err
python - <<'P'
import json
for f in ('gpurun_out/r02_bench_two_stage_2gpu.json','gpurun_out/r02_bench_single_stage_2gpu.json'):
    d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d['parity'], d['config']['exchange'], d['roofline']['kernels_ms_per_step'])
P
