#!/bin/bash
tag=${1:-r2j}
mkdir -p gpurun_out
python tools/bench_phases.py --configs c4x --steps 8 2>gpurun_out/${tag}_c4x.err | tee gpurun_out/${tag}_c4x.json | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'],r['particles'],'loaded elems',r['elements_loaded'],'max ppe',r['max_ppe'],{k:round(v['median_ms'],4) for k,v in r['phases'].items()},'after',r['particles_after'])"
tail -3 gpurun_out/${tag}_c4x.err
ncu --set full --clock-control none --import-source on -k regex:"k_gyro|k_ring|k_scatter" --launch-skip 4 -c 6 -o gpurun_out/${tag}_scatter python tools/bench_phases.py --configs c4x --steps 1 > /dev/null 2>&1
timeout 900 python tools/bench_c3_sweep.py --iters 10 > gpurun_out/${tag}_c3_sweep.jsonl 2> gpurun_out/${tag}_c3_sweep.err; python -c "
import json
for l in open('gpurun_out/${tag}_c3_sweep.jsonl'):
    r=json.loads(l); print(r['series'],r['elements'],r['particles_per_gpu'],r['distribution'],r['op'],round(r['ms_median'],3),'ms',round(r['GBps_at_326B'],1),'GB/s')"
tail -3 gpurun_out/${tag}_c3_sweep.err
