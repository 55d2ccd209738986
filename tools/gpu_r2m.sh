#!/bin/bash
tag=${1:-r2B}
mkdir -p gpurun_out
python tools/bench_picstep.py --steps 30 --timing 2>gpurun_out/${tag}_picstep_1.err | tail -1 | tee -a gpurun_out/${tag}_picstep.jsonl | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print(r['n_gpus'],'ms/step',round(r['ms_per_step'],4),{k:round(v,4) for k,v in r['phase_ms'].items()}); print(json.dumps(r.get('library_phase_avg_ms_rank0'),indent=0))"
python tools/bench_phases.py --configs c2 --steps 12 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'],{k:round(v['median_ms'],4) for k,v in r['phases'].items()})"
