cd /root/repo
for reps in 20 200 3000; do echo "REPS=$reps"; REPS=$reps DEPTHS=3,4,6 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -3; done
for c in 259 333; do echo "CTAS=$c"; TSPWS_CS2_CTAS=$c REPS=3000 DEPTHS=3,4,6 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -3; done
