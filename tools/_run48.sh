cd /root/repo
for c in 25 50 75 100; do
echo "carveout $c morlet"; TSPWS_CS2_CARVEOUT=$c REPS=300 DEPTHS=1,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
echo "carveout $c mexhat"; TSPWS_CS2_CARVEOUT=$c MEXHAT=1 REPS=300 DEPTHS=4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -1
done
