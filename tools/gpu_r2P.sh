#!/bin/bash
# slice scheduling decided by "more slices than NON-EMPTY chunks": tests, diagnostic, configs[3] phases, headline unchanged
tag=${1:-r2P}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_walk_kernels_gpu.py tests/test_search_gpu.py tests/test_structures_gpu.py tests/test_xgc_gpu.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log | cut -c1-300
timeout 500 python tools/diag_c4x.py 2>&1 | tail -4 | cut -c1-330 | tee gpurun_out/${tag}_diag_c4x.txt
timeout 400 python tools/bench_phases.py --configs c4x --steps 8 2>gpurun_out/${tag}_c4x.err | tee gpurun_out/${tag}_c4_xgc2M_phases.json | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'],{k:round(v['median_ms'],4) for k,v in r['phases'].items()}, 'full step', round(r['full_step_ms'],3), r.get('scatter_conservation_full_size'))"
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-picstep 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('headline', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],4), 'ms frac', round(d['roofline']['frac'],3))"
