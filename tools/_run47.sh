cd /root/repo
echo morlet; REPS=300 DEPTHS=1,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
echo mexhat; MEXHAT=1 REPS=300 DEPTHS=1,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
echo "morlet stop after post (TSPWS_DEBUG_STOP=2: tail = group_post2 only)"; TSPWS_DEBUG_STOP=2 REPS=300 DEPTHS=1,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
echo "morlet stop=3 (post + fwd_tail)"; TSPWS_DEBUG_STOP=3 REPS=300 DEPTHS=1,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
