#!/bin/bash
tag=${1:-r2J}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -5 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/bench_picstep.py --steps 20 --timing 2>gpurun_out/${tag}_picstep_2.err | tail -1 | tee -a gpurun_out/${tag}_picstep.jsonl | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print(r['n_gpus'],'ms/step',round(r['ms_per_step'],4),{k:round(v,4) for k,v in r['phase_ms'].items()},'migrated',r['migrated_per_step']); print(r.get('library_phase_avg_ms_rank0'))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 tools/bench_c3_sweep.py --quick --iters 6 > gpurun_out/${tag}_c3_migrate.jsonl 2> gpurun_out/${tag}_c3_migrate.err; python -c "
import json
for l in open('gpurun_out/${tag}_c3_migrate.jsonl'):
    r=json.loads(l); print(r['series'],r['elements'],r['particles_per_gpu'],r['distribution'],r['op'],r['n_gpus'],round(r['ms_median'],3),'ms',round(r['GBps_at_326B'],1),'GB/s deferred',r['deferred'],r['transport'])"
tail -3 gpurun_out/${tag}_c3_migrate.err
