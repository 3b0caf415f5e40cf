cd /root/repo
timeout 300 python tools/_dense_check.py 512 2>&1 | grep -E "dense vs|worst"
timeout 300 python tools/_dense_check.py 2048 2>&1 | grep -E "dense vs|worst"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k two_independent 2>&1 | grep -E "assert|Error|nerr|where" | head
