cd /root/repo
timeout 300 python tools/_dense_check2.py 2>&1 | tail -46
