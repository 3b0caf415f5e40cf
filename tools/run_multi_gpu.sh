#!/bin/bash
# bench.py lines at N GPUs for profiles/ (run through gpurun --gpus N):  bash tools/run_multi_gpu.sh N [out_dir]
cd "$(dirname "$0")/.."
N=${1:-8}
OUT=${2:-gpurun_out}
R=${ROUND:-r02}
mkdir -p "$OUT"
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 20 --warmup 5                                   > "$OUT/${R}_bench_two_stage_${N}gpu.json"    2> "$OUT/b${N}.err";  tail -2 "$OUT/b${N}.err"
run --steps 5 --workload single_stage_w6 --no-strong        > "$OUT/${R}_bench_single_stage_${N}gpu.json" 2> "$OUT/b${N}s.err"; tail -2 "$OUT/b${N}s.err"
run --workload pairs_c5 --steps 8                           > "$OUT/${R}_bench_pairs_${N}gpu.json"        2> "$OUT/b${N}p.err"; tail -2 "$OUT/b${N}p.err"
run --workload resample_c4                                  > "$OUT/${R}_bench_resample_${N}gpu.json"     2> "$OUT/b${N}r.err"; tail -2 "$OUT/b${N}r.err"
python - <<P
import json
N, O, R = $N, "$OUT", "$R"
def last(f): return json.loads(open(f).read().strip().splitlines()[-1])
d = last(f'{O}/{R}_bench_two_stage_{N}gpu.json'); print('two-stage', d['value'], d['ms_per_step'], d['parity']); print(' pipelined', d['config'].get('pipelined')); print(' single', d['config'].get('single_stack')); print(' strong', json.dumps(d['strong'])[:700]); print(' e2e', d['e2e'])
d = last(f'{O}/{R}_bench_single_stage_{N}gpu.json'); print('single-stage', d['value'], d['ms_per_step'], d['parity'])
d = last(f'{O}/{R}_bench_pairs_{N}gpu.json'); print('pairs', d['value'], d.get('pairs_per_s'), d['e2e'])
d = last(f'{O}/{R}_bench_resample_{N}gpu.json'); print('resample', d.get('single_stage'), d.get('two_stage'))
P
