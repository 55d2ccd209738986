#!/bin/bash
# full-size record-multiset check of the PIC step (1 rank; 2 ranks sharing the GPU), memcheck of the new layout kernels
tag=${1:-r2L}
mkdir -p gpurun_out
show() { python -c "
import sys,json
r=json.loads(sys.stdin.read()); print(r['n_gpus'],'ranks ms/step',round(r['ms_per_step'],4),'migrated',r['migrated_per_step'],r.get('full_size_check'))"; }
timeout 300 python tools/bench_picstep.py --steps 10 2>gpurun_out/${tag}_picstep_1.err | tail -1 | tee gpurun_out/${tag}_picstep_1.jsonl | show
tail -3 gpurun_out/${tag}_picstep_1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 tools/bench_picstep.py --shared-gpu --steps 4 --warmup 1 --cube-per-gpu 40 2>gpurun_out/${tag}_picstep_2shared.err | tail -1 | tee gpurun_out/${tag}_picstep_2shared.jsonl | show
tail -3 gpurun_out/${tag}_picstep_2shared.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_structures_gpu.py -m gpu -x -q -k "mostly_empty and 40000" > gpurun_out/${tag}_memcheck_split_rows.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/${tag}_memcheck_split_rows.log | cut -c1-200
