cd /root/repo
timeout 300 python tools/_mma_acc_check2.py 2>&1 | tail -12
