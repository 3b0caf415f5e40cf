cd /root/repo
timeout 900 python -m pytest tests/test_gpu_resample.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/c4_probe.py 2 2>&1 | tail -13
