cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tools/c4_probe.py 2 2>&1 | tail -14
TSPWS_NO_MMA=1 timeout 300 python tools/c4_probe.py 0 2>&1 | grep "single-stage M"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:resample_mma --launch-skip 1 --launch-count 1 -o gpurun_out/r02_resample_mma python tools/c4_probe.py 0 > /dev/null 2>&1
ls -la gpurun_out/r02_resample_mma.ncu-rep
