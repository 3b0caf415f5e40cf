cd /root/repo
TSPWS_CS3=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_stage or pipeline or peer_exchange" 2>&1 | tail -3
echo "single stack CS3=1"; TSPWS_CS3=1 PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -2
echo "single stack CS3=0"; TSPWS_CS3=0 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
echo "pipe CS3 default"; REPS=200 DEPTHS=1,3,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -3
echo "pipe CS3=0"; TSPWS_CS3=0 REPS=200 DEPTHS=3,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
