#!/bin/bash
# two ranks sharing one GPU (hosted bootstrap, no NCCL) + L2 window A/B
tag=${1:-r2F}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -k "single_rank or one_gpu" > gpurun_out/${tag}_shared_gpu_tests.log 2>&1; tail -15 gpurun_out/${tag}_shared_gpu_tests.log | cut -c1-600
out=gpurun_out/${tag}_l2_window_ab.txt
: > $out
line() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],4), 'ms frac', round(d['roofline']['frac'],3))"; }
for w in 0 0.5 1.0; do
  for g in "" "--no-graph"; do
    PUMIPIC_L2_WINDOW=$w timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-picstep $g 2>>gpurun_out/${tag}_bench.err | tail -1 | line "window=$w $g" | tee -a $out
  done
done
for w in 0 1.0; do
  PUMIPIC_L2_WINDOW=$w timeout 300 python tools/bench_picstep.py --steps 20 2>>gpurun_out/${tag}_picstep.err | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('window=$w picstep ms/step',round(r['ms_per_step'],4),{k:round(v,4) for k,v in r['phase_ms'].items()})" | tee -a $out
done
