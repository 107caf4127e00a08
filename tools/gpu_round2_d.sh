#!/bin/bash
# round 2, call D (N GPUs): the driver's scaling command for C4 (strong by default, weak measured in the same run)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/d_bench_n$N.json 2> gpurun_out/d_bench_n$N.err
echo "rc $?"; cat gpurun_out/d_bench_n$N.json; tail -15 gpurun_out/d_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/h2d_bw.py 2>/dev/null | tail -1 | tee gpurun_out/d_h2d_n$N.json
