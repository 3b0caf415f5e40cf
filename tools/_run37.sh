cd /root/repo
timeout 300 python tools/pipe_probe.py 2>&1 | tail -5
