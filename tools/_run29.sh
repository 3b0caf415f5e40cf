cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer or distributed" 2>&1 | tail -5
