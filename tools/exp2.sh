#!/bin/bash
for v in 0 1 2 3 4; do
  LHG_BLOCK_COLS_LOG2=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/exp_b$v.json 2>gpurun_out/exp_b$v.err
  echo "block_cols_log2=$v"; python tools/bsum.py gpurun_out/exp_b$v.json
done
