#!/bin/bash
# run the C4 bench once per experiment library (tools/: diagnosis only, numbers are not bench values)
for v in "$@"; do
  if [ "$v" = "base" ]; then lib=""; else lib="$PWD/learned_hologram_gan_b200/lib/libasm_b200_$v.so"; fi
  LHG_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/exp_$v.json 2>gpurun_out/exp_$v.err
  echo "variant=[$v]"; python tools/bsum.py gpurun_out/exp_$v.json
done
