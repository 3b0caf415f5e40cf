cd /root/repo
echo "pad default (on in pipe)"; REPS=300 DEPTHS=1,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
echo "mexhat"; MEXHAT=1 REPS=300 DEPTHS=1,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
echo "pad off"; TSPWS_CS2_PAD=0 REPS=300 DEPTHS=4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -1
for b in 16384 32768 65536; do echo "pad bytes $b"; TSPWS_CS2_PAD_BYTES=$b REPS=300 DEPTHS=4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -1;  done
echo "single stack with pad"; TSPWS_CS2_PAD=1 timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
echo "single stack without"; timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
