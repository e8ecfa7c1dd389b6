#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/peer_gather_check.py > gpurun_out/peer_check.log 2>&1
echo "peer check exit $?"; grep -E "world|Error|error|warn" gpurun_out/peer_check.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "bench n$N exit $?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-260
