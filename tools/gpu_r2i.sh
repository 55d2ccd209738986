#!/bin/bash
tag=${1:-r2i}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${tag}_pytest_gpu.log | cut -c1-400
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python -c "
import json; r=json.load(open('gpurun_out/${tag}_bench.json'))
print('value',r['value'],'ms',r['ms_per_step'],'frac',r['roofline']['frac']); print('e2e',r['e2e']['value'],r['e2e']['h2d_bytes_per_step']); print('cpu',r['cpu_baseline']['value'],r['cpu_baseline']['cores'],r['cpu_baseline']['ms_per_step']); print('parity',r['parity']); p=r['picstep']; print('picstep',p['ms_per_step'],p['phase_ms'],p['parity'])"
tail -3 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_bench_reference.json
