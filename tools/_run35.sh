set -x
cd /root/repo
for d in 16 8 4; do
TSPWS_BAND_D=$d PROF=1 timeout 300 python tools/quick_bench.py c2 2>&1 | tail -2
done
TSPWS_BAND_D=8 timeout 600 python tools/_band_check.py 1024 16384 2>&1 | grep -v "dense only" | head -14
