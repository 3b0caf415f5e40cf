cd /root/repo
for m in 1 3 5 7; do echo "terms $m"; TSPWS_DENSE_TERMS=$m timeout 300 python tools/_dense_check.py 512 2>&1 | grep -E "dense vs"; done
