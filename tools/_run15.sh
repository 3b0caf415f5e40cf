cd /root/repo
timeout 120 python -m pytest tests/test_gpu_mma.py -m gpu -x -q 2>&1 | tail -15
