#!/bin/bash
tag=${1:-r2g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
run() { python tools/bench_phases.py --steps 12 "$@" 2>>gpurun_out/${tag}_phases.err | tee -a gpurun_out/${tag}_phases.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'],'mode',r['rebuild_mode'],'shuf',r['shuffling'],{k:round(v['median_ms'],4) for k,v in r['phases'].items()})"; }
run --configs c2 --rebuild-mode 2
run --configs c2 --rebuild-mode 2 --shuffling 0
run --configs c2 --rebuild-mode 1
run --configs c4 --rebuild-mode 2
run --configs c4 --rebuild-mode 1
run --configs c3 --rebuild-mode 2
