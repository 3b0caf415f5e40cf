cd /root/repo
export TSPWS_FUSED_COMBINE=0
for c in 592 1184 1776 2368 592 1184; do
echo "== CS2_CTAS=$c"; TSPWS_CS2_CTAS=$c timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
done
TSPWS_CS2_CTAS=1184 PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -2
