cd /root/repo
timeout 300 python tools/_mma_acc_check.py 2>&1 | tail -10
