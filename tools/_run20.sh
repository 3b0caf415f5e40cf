cd /root/repo
timeout 300 python tools/_dense_check.py 512 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -4
echo "== dense on"; PROF=1 timeout 300 python tools/quick_bench.py 2>&1 | tail -2
timeout 600 python bench.py --workload single_stage_w6 --steps 3 --no-cpu --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernels_ms_per_step'])"
