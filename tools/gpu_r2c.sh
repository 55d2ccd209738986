#!/bin/bash
tag=${1:-r2c}
mkdir -p gpurun_out
python -m pytest tests/test_structures_gpu.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -5 gpurun_out/${tag}_pytest.log
run() { python tools/bench_phases.py --configs c2 --steps 12 --shuffling 0 "$@" 2>>gpurun_out/${tag}_phases.err | tee -a gpurun_out/${tag}_phases.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('mode',r['rebuild_mode'],'order',r['chunk_order'],'tuning',r['tuning'],'rebuild ms',round(r['phases']['rebuild']['median_ms'],4),'min',round(r['phases']['rebuild']['min_ms'],4))"; }
run --rebuild-mode 1
run --rebuild-mode 2
run --rebuild-mode 2 --chunk-order 0
run --rebuild-mode 2 --tuning 2,0,-1
run --rebuild-mode 2 --tuning 8,0,-1
run --rebuild-mode 2 --tuning 4,0,1
run --rebuild-mode 3 --tuning 0,1,-1
run --rebuild-mode 3 --tuning 0,2,-1
run --rebuild-mode 3 --tuning 0,4,-1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv python tools/bench_phases.py --configs c2 --steps 2 --shuffling 0 > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.txt 2>/dev/null; head -12 gpurun_out/${tag}_launches.txt
