#!/bin/bash
tag=${1:-r2h}
mkdir -p gpurun_out
run() { python tools/bench_phases.py --steps 10 "$@" 2>>gpurun_out/${tag}_phases.err | tee -a gpurun_out/${tag}_phases.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'],'mode',r['rebuild_mode'],'tuning',r['tuning'],{k:round(v['median_ms'],4) for k,v in r['phases'].items() if 'rebuild' in k})"; }
run --configs c2 --rebuild-mode 2
for t in 0 1 2 3 5; do run --configs c4 --rebuild-mode 2 --tuning $t; done
