cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in 0 1 2 3; do
  echo "== TSPWS_FW_FLAGS=$f fused=0"
  TSPWS_FW_FLAGS=$f TSPWS_FUSED_COMBINE=0 PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -2
  echo "== TSPWS_FW_FLAGS=$f fused=1"
  TSPWS_FW_FLAGS=$f TSPWS_FUSED_COMBINE=1 PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -2
done
for f in 0 1 3; do
  echo "== ss TSPWS_FW_FLAGS=$f"
  TSPWS_FW_FLAGS=$f timeout 300 python tools/quick_bench.py ss 2>&1 | tail -1
done
