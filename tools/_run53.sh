cd /root/repo
timeout 300 python tools/pipe_timeline.py 2>&1 | grep -A3 "records\|colsum2 launches\|fwd_tail launches\|inv4 launches"
for c in 296 444 592; do echo "CTAS=$c"; TSPWS_CS2_CTAS=$c REPS=300 DEPTHS=1,3,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -3; done
echo mexhat; MEXHAT=1 REPS=300 DEPTHS=4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -1
echo prio=tail; TSPWS_PIPE_PRIO=tail REPS=300 DEPTHS=4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -1
