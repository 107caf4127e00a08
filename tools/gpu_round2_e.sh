#!/bin/bash
# round 2, call E (N GPUs): every workload under the driver's torchrun command line, plus the host-link ceiling
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=${1:-8}
mkdir -p gpurun_out
run() {  # workload, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $1 --steps ${2:-10} --warmup 3 > gpurun_out/e_bench_$1_n$N.json 2> gpurun_out/e_bench_$1_n$N.err
  echo "$1 rc $?"; head -c 1500 gpurun_out/e_bench_$1_n$N.json; echo; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/e_bench_$1_n$N.err | tail -4
}
run c4 10
run c5 5
run c3 20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/h2d_bw.py 2>/dev/null | tail -1 | tee gpurun_out/e_h2d_n$N.json
