cd /root/repo
for pr in equal sums tail; do
echo "2 tail streams, prio=$pr"; TSPWS_PIPE_TAILS=2 TSPWS_PIPE_PRIO=$pr REPS=400 DEPTHS=3,4,6 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -3
done
TSPWS_PIPE_TAILS=2 TSPWS_PIPE_PRIO=equal OUT=gpurun_out/tl_2t.txt timeout 300 python tools/pipe_timeline.py 2>&1 | grep records
python tools/tl_seq.py gpurun_out/tl_2t.txt 1000 1400
