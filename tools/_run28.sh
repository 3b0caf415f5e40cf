cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fwd_dense_mma --launch-skip 1 --launch-count 1 -o gpurun_out/r02_fwd_dense_mma python bench.py --workload single_stage_w6 --steps 2 --no-cpu --no-extra > /dev/null 2>&1
ls -la gpurun_out/r02_fwd_dense_mma.ncu-rep
timeout 600 python bench.py --workload single_stage_w6 --steps 5 --no-extra 2>/dev/null > gpurun_out/r02_bench_single_stage.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_single_stage.json')); print(d['ms_per_step'], d['value'], d['cpu_baseline'], d['parity'])"
