cd /root/repo
time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -22
