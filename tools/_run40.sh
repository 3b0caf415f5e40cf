cd /root/repo
for c in 592 444 370 296 222; do echo "CS2_CTAS=$c"; TSPWS_CS2_CTAS=$c timeout 300 python tools/pipe_probe3.py 2>&1 | tail -4; done
