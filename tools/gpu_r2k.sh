#!/bin/bash
tag=${1:-r2k}
mkdir -p gpurun_out
python -m pytest tests/test_walk_kernels_gpu.py tests/test_search_gpu.py tests/test_structures_gpu.py tests/test_xgc_gpu.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log | cut -c1-300
python bench.py --no-cpu-baseline --no-e2e --no-picstep 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],4), 'ms frac', round(d['roofline']['frac'],3))"
python tools/bench_phases.py --configs c4x --steps 8 2>gpurun_out/${tag}_c4x.err | tee gpurun_out/${tag}_c4x.json | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'],r['particles'],'loaded elems',r['elements_loaded'],'max ppe',r['max_ppe'],{k:round(v['median_ms'],4) for k,v in r['phases'].items()},'after',r['particles_after'])"
tail -3 gpurun_out/${tag}_c4x.err
