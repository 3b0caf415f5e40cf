cd /root/repo
for ch in 592 1184 2368; do for c in 592 444 296; do echo "CHUNKS=$ch CTAS=$c"; TSPWS_CS2_CHUNKS=$ch TSPWS_CS2_CTAS=$c timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1; done; done
for ch in 592 1184 2368; do echo "pipe CHUNKS=$ch"; TSPWS_CS2_CHUNKS=$ch timeout 300 python tools/pipe_probe3.py 2>&1 | tail -4; done
echo "pipe CHUNKS=1184 CTAS=444"; TSPWS_CS2_CTAS=444 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -4
