#!/bin/bash
# block-level destination histogram: parity tests, A/B on C2 and on configs[3] (xgc/2M.osh), scatter conservation at full size
tag=${1:-r2O}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_structures_gpu.py tests/test_xgc_gpu.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log | cut -c1-300
show() { python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'],'hist',r['block_histogram'],{k:round(v['median_ms'],4) for k,v in r['phases'].items()}, r.get('scatter_conservation_full_size'))"; }
for h in 1 0; do
  timeout 300 python tools/bench_phases.py --configs c2 --steps 12 --block-histogram $h 2>/dev/null | tee -a gpurun_out/${tag}_phases.jsonl | show
done
for h in 1 0; do
  timeout 400 python tools/bench_phases.py --configs c4x --steps 8 --block-histogram $h 2>gpurun_out/${tag}_c4x_$h.err | tee -a gpurun_out/${tag}_phases.jsonl | show
done
tail -2 gpurun_out/${tag}_c4x_1.err
