#!/bin/bash
# round 2, call P: split-phase column kernel + warp-local row passes as the default build -- full -m gpu suite,
# A/B of the warp-local K1/K3 (variant nok13 = without), C5 line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/p_pytest.log; tail -3 gpurun_out/p_pytest.log
bash tools/exp.sh nok13 base nok13 base
timeout 900 python bench.py --workload c5 --steps 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/p_bench_c5.json 2> gpurun_out/p_bench_c5.err; echo "c5 rc $?"
LHG_LIB=$PWD/learned_hologram_gan_b200/lib/libasm_b200_nok13.so timeout 900 python bench.py --workload c5 --steps 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/p_bench_c5_nok13.json 2> gpurun_out/p_bench_c5_nok13.err; echo "c5 rc $?"
python - <<'PY'
import json
for f in ("p_bench_c5", "p_bench_c5_nok13"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]), round(d["ms_per_step"], 2), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["per_kernel"].items()})
    except Exception as e: print(f, "ERR", e)
PY
