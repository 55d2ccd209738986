#!/bin/bash
tag=${1:-r2D}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -12 gpurun_out/${tag}_tests.log | cut -c1-400
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/bench_picstep.py --steps 20 --timing "$@" 2>gpurun_out/${tag}_picstep_$n.err | tail -1 | tee -a gpurun_out/${tag}_picstep.jsonl | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print(r['n_gpus'],'ms/step',round(r['ms_per_step'],4),{k:round(v,4) for k,v in r['phase_ms'].items()},'migrated',r['migrated_per_step'])"; }
run 2
PUMIPIC_P2P=0 run 2
