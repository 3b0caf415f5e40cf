set -x
cd /root/repo
timeout 600 python tools/_band_check.py 1024 16384 2>&1 | tail -60
PROF=1 timeout 300 python tools/quick_bench.py c2 2>&1 | tail -4
PROF=1 TSPWS_BAND_D=0 timeout 300 python tools/quick_bench.py c2 2>&1 | tail -4
