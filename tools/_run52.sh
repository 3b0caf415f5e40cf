cd /root/repo
timeout 300 python tools/pipe_timeline.py 2>&1 | tail -70
