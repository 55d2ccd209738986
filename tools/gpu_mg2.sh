#!/bin/bash
# 2-GPU call: multi-rank parity worker, balancer scenario, picstep at 1 and 2 ranks, bench.py --gpus 2
tag=${1:-r2m}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/${tag}_multigpu.log 2>&1; tail -5 gpurun_out/${tag}_multigpu.log
MGPU_ONLY=balancer timeout -s USR1 -k 15 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/mgpu_worker.py > gpurun_out/${tag}_balancer.log 2>&1; echo "balancer rc=$?"; tail -25 gpurun_out/${tag}_balancer.log
for n in 1 2; do
 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/bench_picstep.py --steps 10 2>gpurun_out/${tag}_picstep_$n.err | tail -1 | tee -a gpurun_out/${tag}_picstep.jsonl | cut -c1-900
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/${tag}_bench2.err | tail -1 | tee gpurun_out/${tag}_bench2.json | cut -c1-2500
tail -5 gpurun_out/${tag}_bench2.err
