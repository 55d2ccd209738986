#!/bin/bash
# 2-GPU call: multi-rank parity worker over both transports, picstep at 1 and 2 ranks, bench.py --gpus 2
tag=${1:-r2m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/${tag}_multigpu.log 2>&1; tail -30 gpurun_out/${tag}_multigpu.log | cut -c1-400
for n in 1 2; do
 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/bench_picstep.py --steps 10 2>gpurun_out/${tag}_picstep_$n.err | tail -1 | tee -a gpurun_out/${tag}_picstep.jsonl | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print(r['n_gpus'],'ms/step',round(r['ms_per_step'],4),{k:round(v,4) for k,v in r['phase_ms'].items()},'migrated',r['migrated_per_step'])"
done
PUMIPIC_P2P=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 tools/bench_picstep.py --steps 10 2>gpurun_out/${tag}_picstep_2nccl.err | tail -1 | tee -a gpurun_out/${tag}_picstep.jsonl | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('nccl',r['n_gpus'],'ms/step',round(r['ms_per_step'],4),{k:round(v,4) for k,v in r['phase_ms'].items()},'migrated',r['migrated_per_step'])"
tail -3 gpurun_out/${tag}_picstep_2.err
