cd /root/repo
TSPWS_CS3=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_stage or pipeline or peer_exchange" 2>&1 | tail -3
for v in "" tools/_v/r1s24 tools/_v/r4s6 tools/_v/r2s16 tools/_v/r2s8; do
  L=""; if [ -n "$v" ]; then L=/root/repo/$v/libtspws_b200.so; fi
  echo "== variant ${v:-r2s12}"
  TSPWS_LIB=$L TSPWS_CS3=1 PROF=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -2 | cut -c1-160
  TSPWS_LIB=$L REPS=200 DEPTHS=4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -1
done
