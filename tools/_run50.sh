cd /root/repo
for c in 148 176 208 296; do
echo "CTAS=$c full"; TSPWS_CS2_CTAS=$c REPS=300 DEPTHS=1,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
echo "CTAS=$c stop=3"; TSPWS_DEBUG_STOP=3 TSPWS_CS2_CTAS=$c REPS=300 DEPTHS=1,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
echo "CTAS=$c stop=2"; TSPWS_DEBUG_STOP=2 TSPWS_CS2_CTAS=$c REPS=300 DEPTHS=4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -1
done
