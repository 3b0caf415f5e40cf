#!/bin/bash
# rebuild the CUDA TU with different forward-kernel launch bounds and time both paths (run on the GPU box)
cd "$(dirname "$0")/../ts-pws_b200/csrc"
for cfg in "-DFW_WARPS=4 -DFW_MINB=4" "-DFW_WARPS=4 -DFW_MINB=5" "-DFW_WARPS=4 -DFW_MINB=6" "-DFW_WARPS=8 -DFW_MINB=2" "-DFW_WARPS=2 -DFW_MINB=8" "-DFW_WARPS=4 -DFW_MINB=3"; do
  rm -f tspws_kernels.o libtspws_b200.so
  make EXTRA="$cfg" > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  echo "== $cfg  $(grep -A2 'fwd2_pertrace' tspws_kernels.ptxas.log | grep -oE 'Used [0-9]+ registers' | head -1)"
  (cd ../..; python tools/quick_bench.py ts | tail -1; python tools/quick_bench.py ss | tail -1)
done
rm -f tspws_kernels.o libtspws_b200.so; make > /dev/null 2>&1
