set -x
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -3
timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
TSPWS_FUSED_COMBINE=0 PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -3
TSPWS_FUSED_COMBINE=0 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
timeout 300 python tools/quick_bench.py ss 2>&1 | tail -1
