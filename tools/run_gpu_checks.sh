#!/bin/bash
# Reproduces the round-2 files under profiles/ on a B200 box (run from the repo root, e.g. through gpurun):
#   bash tools/run_gpu_checks.sh [out_dir]
# 1 GPU is enough; the multi-GPU lines need  gpurun --gpus N  and the torchrun command at the end.
set -u
OUT=${1:-gpurun_out}
R=${ROUND:-r02}
mkdir -p "$OUT"
T="timeout 900"

$T python -m pytest tests -m gpu -q 2>&1 | tail -3
$T python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

# bench lines (one JSON line each on stdout)
$T python bench.py                                       > "$OUT/${R}_bench_two_stage.json"     2> "$OUT/bench_two_stage.err"
$T python bench.py --workload single_stage_w6 --steps 5  > "$OUT/${R}_bench_single_stage.json"  2> "$OUT/bench_single.err"
$T python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/${R}_bench_reference_arm.json" 2> "$OUT/bench_ref.err"

# launch lists (cache-flushed, serialised: compare shares, never quote these times as bench values)
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/${R}_launches_two_stage.csv" \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${R}_launches_single_stage.csv" \
    python bench.py --workload single_stage_w6 --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1

# full metric sets of the dominant kernel and of the tail kernels (warm caches), read back with
#   ncu -i X.ncu-rep --page raw --csv > profiles/Y.csv ; python tools/make_profile_tables.py
$T ncu --set full --import-source on --clock-control none --cache-control none -k regex:"colsum2" \
    --launch-skip 2 --launch-count 1 -o "$OUT/${R}_colsum2_full" python tools/quick_bench.py ts > /dev/null 2>&1
$T ncu --set full --import-source on --clock-control none --cache-control none \
    -k regex:"group_post2|fwd_tail|phasor_combine|inv2|inv4|inv_finish" --launch-skip 12 --launch-count 6 \
    -o "$OUT/${R}_tail_full" python tools/quick_bench.py ts > /dev/null 2>&1
# the tcgen05 kernels of the single-stage forward transform (dense + banded GEMM)
$T ncu --set full --import-source on --clock-control none -k regex:"fwd_mma_persistent|fwd_dense_mma" --launch-skip 1 --launch-count 1 \
    -o "$OUT/${R}_fwd_band_mma" python tools/quick_bench.py c2 > /dev/null 2>&1

if [ -n "${EXTRA_PROBES:-}" ]; then
# stage budget of the two-stage step (graph replay prefixes) and the design micro-benchmarks
for s in 1 2 3 4 0; do echo "TSPWS_DEBUG_STOP=$s"; TSPWS_DEBUG_STOP=$s python tools/quick_bench.py ts | tail -1; done
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/readbw tools/readbw.cu 2>/dev/null && tools/readbw
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/graph_chain tools/graph_chain.cu 2>/dev/null && tools/graph_chain
# configs[3] / configs[4] probes, the stack pipeline
$T python tools/c4_probe.py 2
$T python tools/batch_bench.py 64 0
$T python tools/batch_bench.py 64 10
REPS=300 DEPTHS=1,3,4 $T python tools/pipe_probe.py
# sanitizers
$T compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q 2>&1 | tail -3
fi

cat <<'EON'
multi-GPU (needs N GPUs):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus N --steps 50 --warmup 3            # peer exchange; TSPWS_BENCH_NCCL=1 for the NCCL path
  python -m pytest tests/test_distributed.py -m gpu -q   # needs 2 GPUs
per-CTA timeline of the stack pipeline (diagnostics build of the library):
  mkdir -p tools/_tl && cp ts-pws_b200/csrc/{*.c,*.cu,*.h,Makefile} tools/_tl/ && make -C tools/_tl EXTRA=-DTSPWS_TIMELINE_BUILD libtspws_b200.so
  python tools/pipe_timeline.py && python tools/tl_seq.py gpurun_out/pipe_timeline.txt 1000 1500
EON
