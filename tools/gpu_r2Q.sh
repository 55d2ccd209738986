#!/bin/bash
# last validation of the round: whole GPU suite, smoke, the bench line, ps_combo160 rebuild quick sweep
tag=${1:-r2Q}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2>gpurun_out/${tag}_bench.err; python -c "
import json; r=json.loads(open('gpurun_out/${tag}_bench_n1.json').read().strip().splitlines()[-1])
print('value',r['value'],'ms',r['ms_per_step'],'frac',r['roofline']['frac'],'e2e',r['e2e']['value'],'cpu',r['cpu_baseline']['value'],'parity',r['parity']['mismatch']); p=r['picstep']; print('picstep',p['ms_per_step'],p['phase_ms'],p.get('parity',{}).get('mismatch'),{k:v for k,v in p.get('full_size_check',{}).items() if k!='what'})"
timeout 400 python tools/bench_c3_sweep.py --quick --iters 5 2>gpurun_out/${tag}_c3.err | tee gpurun_out/${tag}_c3_sweep_rebuild_quick.jsonl | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): continue
    r=json.loads(l); print(r['series'],r['elements'],r['particles_per_gpu'],r['distribution'],r['op'],round(r['ms_median'],3),'ms',round(r['GBps_at_326B'],1),'GB/s')"
