cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_stage or golden or medium" 2>&1 | tail -6
echo "== dense on"; PROF=1 timeout 300 python tools/quick_bench.py 2>&1 | tail -2
echo "== dense off"; TSPWS_DENSE_D=0 PROF=1 timeout 300 python tools/quick_bench.py 2>&1 | tail -2
