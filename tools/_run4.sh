cd /root/repo
TSPWS_TIMELINE=gpurun_out/tl_f1.txt TSPWS_NO_GRAPH=1 TSPWS_FW_FLAGS=1 TSPWS_FUSED_COMBINE=0 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
TSPWS_TIMELINE=gpurun_out/tl_f1c.txt TSPWS_NO_GRAPH=1 TSPWS_FW_FLAGS=1 TSPWS_FUSED_COMBINE=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
TSPWS_TIMELINE=gpurun_out/tl_f0.txt TSPWS_NO_GRAPH=1 TSPWS_FW_FLAGS=0 TSPWS_FUSED_COMBINE=0 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
