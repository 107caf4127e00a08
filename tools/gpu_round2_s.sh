#!/bin/bash
# round 2, call S: do the clock sampler's NVML calls stretch steps of the timed region?  (diagnosis)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for mode in nvml off nvml off nvml off; do
  for poll in 25; do
    LHG_CLOCK_SAMPLER=$mode LHG_CLOCK_POLL_MS=$poll python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/s_$mode.json 2>gpurun_out/s_$mode.err
    echo "$mode $poll $(python tools/bsum.py gpurun_out/s_$mode.json | cut -c1-60)"
  done
done
for poll in 100 100 100; do
    LHG_CLOCK_POLL_MS=$poll python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/s_p.json 2>gpurun_out/s_p.err
    echo "nvml $poll $(python tools/bsum.py gpurun_out/s_p.json | cut -c1-60)"
done
