set -x
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -4
timeout 300 python tools/quick_bench.py 2>&1 | tail -5
timeout 600 ncu --set full --import-source on --clock-control none -k regex:select_sum --launch-skip 2 --launch-count 2 -o gpurun_out/r02_select_sum python tools/select_probe.py 256 128 2>&1 | tail -5
