#!/bin/bash
# split-rows layout + record-cooperative pack: parity tests, then the sparse-structure rebuild timing
tag=${1:-r2I}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_structures_gpu.py -x -q > gpurun_out/${tag}_structures.log 2>&1; tail -5 gpurun_out/${tag}_structures.log | cut -c1-400
timeout 900 python -m pytest tests/test_multigpu.py -x -q -k "single_rank or one_gpu" > gpurun_out/${tag}_shared_gpu_tests.log 2>&1; tail -8 gpurun_out/${tag}_shared_gpu_tests.log | cut -c1-600
timeout 600 python tools/bench_sparse_rebuild.py 2>gpurun_out/${tag}_sparse.err | tee gpurun_out/${tag}_sparse_rebuild.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('split',r['split_rows'],'rebuild ms',round(r['rebuild_ms_median'],4),r['library_phase_avg_ms'])"
tail -3 gpurun_out/${tag}_sparse.err
timeout 300 python tools/bench_picstep.py --steps 20 --timing 2>>gpurun_out/${tag}_picstep.err | tail -1 | tee -a gpurun_out/${tag}_picstep_1.jsonl | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('picstep ms/step',round(r['ms_per_step'],4),{k:round(v,4) for k,v in r['phase_ms'].items()})"
