#!/bin/bash
for v in "2 1" "2 2" "1 1" "3 1" "2 0"; do
  set -- $v
  LHG_BLOCK_W1=$1 LHG_BLOCK_W2=$2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/exp_l.json 2>gpurun_out/exp_l.err
  echo "W1 block=$1 W2 block=$2"; python tools/bsum.py gpurun_out/exp_l.json
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fast_kernel|col_warp" -s 18 -c 6 --csv --log-file gpurun_out/step_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
