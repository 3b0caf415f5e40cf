cd /root/repo
for c in 592 444 296 148; do echo "CS2_CTAS=$c"; TSPWS_CS2_CTAS=$c timeout 300 python tools/pipe_probe.py 2>&1 | tail -3; done
