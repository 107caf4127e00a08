#!/bin/bash
# round 2, call W (N GPUs): the C4 line (strong split + weak rate in one run) of the final build under the driver's
# torchrun command line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/w_bench_c4_n$N.json 2> gpurun_out/w_bench_c4_n$N.err
echo "c4 rc $?"; head -c 2500 gpurun_out/w_bench_c4_n$N.json; echo; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/w_bench_c4_n$N.err | tail -4
