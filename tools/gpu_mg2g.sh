#!/bin/bash
tag=${1:-r2w}
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/bench_picstep.py --steps 20 "$@" 2>gpurun_out/${tag}_picstep_$n.err | tail -1 | tee -a gpurun_out/${tag}_picstep.jsonl | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print(r['n_gpus'],r['comm_array_reduce'][:10],'ms/step',round(r['ms_per_step'],4),{k:round(v,4) for k,v in r['phase_ms'].items()},'migrated',r['migrated_per_step']); print(r.get('library_phase_avg_ms_rank0'))"; }
run $N --timing
run $N --inline-reduce
tail -3 gpurun_out/${tag}_picstep_$N.err
