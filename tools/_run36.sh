set -x
cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or full_size_single or large_single" 2>&1 | tail -5
PROF=1 timeout 300 python tools/quick_bench.py c2 2>&1 | tail -2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fwd_dense_mma --launch-skip 2 --launch-count 2 -o gpurun_out/r02_fwd_band_mma python tools/quick_bench.py c2 2>&1 | tail -3
