cd /root/repo
timeout 300 python tools/_dense_check.py 512 2>&1 | tail -6
