#!/bin/bash
# W1/W2 block-width sweep (LHG_BLOCK_W1 / LHG_BLOCK_W2 = log2 of the block columns, 0 = plain) for one workload
wl=${1:-c4}
for v in "0 0" "0 1" "0 2" "2 0" "2 1" "2 2" "1 1"; do
  set -- $v
  LHG_BLOCK_W1=$1 LHG_BLOCK_W2=$2 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp_l.json 2>gpurun_out/exp_l.err
  echo "W1 block=$1 W2 block=$2"; python tools/bsum.py gpurun_out/exp_l.json
done
