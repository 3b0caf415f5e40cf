cd /root/repo
export TSPWS_FUSED_COMBINE=0 TSPWS_FW_FLAGS=1
for o in 0 1 0 1; do
  echo "== order $o"; TSPWS_TAIL_ORDER=$o timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
done
TSPWS_TAIL_ORDER=1 PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -2
TSPWS_TIMELINE=gpurun_out/tl_o1.txt TSPWS_NO_GRAPH=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
