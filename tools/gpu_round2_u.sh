#!/bin/bash
# round 2, call U: A/B of a variant library against the default build on C4 and C5 (+ parity subset on the default)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
V=$1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config4 or full_size or fused or sharded or repeated or config5 or multi_distance or forward_methods or gradients" > gpurun_out/u_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/u_pytest.log
bash tools/exp.sh $V base $V base
for v in $V base; do
  if [ "$v" = "base" ]; then lib=""; else lib="$PWD/learned_hologram_gan_b200/lib/libasm_b200_$v.so"; fi
  LHG_LIB=$lib timeout 900 python bench.py --workload c5 --steps 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/u_c5_$v.json 2> gpurun_out/u_c5_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/u_c5_$v.json")); print("c5 $v", round(d["value"]), round(d["ms_per_step"], 2), {k: round(x["ms_per_step"], 2) for k, x in d["roofline"]["per_kernel"].items()})
except Exception as e: print("c5 $v ERR", e)
PY
done
