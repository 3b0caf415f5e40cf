cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for fc in 0 1; do for o in 0 1; do
  echo "== fused=$fc order=$o"; TSPWS_FUSED_COMBINE=$fc TSPWS_TAIL_ORDER=$o timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
done; done
TSPWS_FUSED_COMBINE=0 PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -2
TSPWS_FUSED_COMBINE=0 TSPWS_TIMELINE=gpurun_out/tl_v2.txt TSPWS_NO_GRAPH=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
timeout 300 python tools/quick_bench.py 2>&1 | tail -4
