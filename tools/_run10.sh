cd /root/repo
export TSPWS_FUSED_COMBINE=0
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
echo "== r1 lib"; TSPWS_LIB=tools/_r1/libtspws_b200.so timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
echo "== new lib"; timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
done
PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -2
TSPWS_NO_GRAPH=1 TSPWS_LIB=tools/_tl/libtspws_b200.so TSPWS_TIMELINE=gpurun_out/tl_all2.txt timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
