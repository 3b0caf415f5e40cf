cd /root/repo
for pr in equal equalhi; do
echo "prio=$pr"; TSPWS_PIPE_PRIO=$pr REPS=300 DEPTHS=3,4 timeout 300 python tools/pipe_probe3.py 2>&1 | tail -2
TSPWS_PIPE_PRIO=$pr OUT=gpurun_out/tl_$pr.txt timeout 300 python tools/pipe_timeline.py 2>&1 | grep records
python tools/tl_seq.py gpurun_out/tl_$pr.txt 1000 1500
done
