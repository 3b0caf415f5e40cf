cd /root/repo
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
( time run 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json ) 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','parity','n_gpus')})
print('strong', json.dumps(d['strong'], indent=1))
print('e2e', d['e2e'])
print('kern', d['roofline']['kernels_ms_per_step'])
P
timeout 600 python -m pytest tests/test_distributed.py -m gpu -x -q 2>&1 | tail -3
( time run 2 --workload pairs_c5 --steps 4 > gpurun_out/bench_pairs_n2.json ) 2> gpurun_out/bench_pairs_n2.err; tail -3 gpurun_out/bench_pairs_n2.err; cat gpurun_out/bench_pairs_n2.json | cut -c1-900
( time run 2 --workload resample_c4 > gpurun_out/bench_res_n2.json ) 2> gpurun_out/bench_res_n2.err; tail -3 gpurun_out/bench_res_n2.err; cat gpurun_out/bench_res_n2.json | cut -c1-1200
