#!/bin/bash
# 8-GPU record of the final code: bench.py --gpus 8 (headline + e2e + picstep with parity)
tag=${1:-r2K}
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/${tag}_bench$N.err | tail -1 > gpurun_out/${tag}_bench$N.json; python -c "
import json; r=json.load(open('gpurun_out/${tag}_bench$N.json'))
print('value',r['value'],'ms',r['ms_per_step'],'e2e',r['e2e']['value']); p=r['picstep']; print('picstep',p['ms_per_step'],p['phase_ms'],p['parity'])"
tail -3 gpurun_out/${tag}_bench$N.err
