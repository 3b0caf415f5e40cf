cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for i in 1 2; do
echo "== r1 lib"; TSPWS_LIB=tools/_r1/libtspws_b200.so timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
echo "== new lib"; timeout 300 python tools/quick_bench.py ts 2>&1 | tail -1
done
