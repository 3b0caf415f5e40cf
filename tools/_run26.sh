cd /root/repo
timeout 300 python tools/_dense_check2.py 2>&1 | grep -E "^dense|^simt|scale 39 D"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_mma.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload single_stage_w6 --steps 3 --no-cpu --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernels_ms_per_step'])"
