#!/bin/bash
tag=${1:-r2f}
mkdir -p gpurun_out
python -m pytest tests/test_structures_gpu.py -m gpu -x -q -k "gather or ordered" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
run() { python tools/bench_phases.py --configs c2 --steps 12 --shuffling 0 "$@" 2>>gpurun_out/${tag}_phases.err | tee -a gpurun_out/${tag}_phases.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('mode',r['rebuild_mode'],'order',r['chunk_order'],'tuning',r['tuning'],'rebuild ms',round(r['phases']['rebuild']['median_ms'],4),'min',round(r['phases']['rebuild']['min_ms'],4))"; }
for b in 3 5 7; do run --rebuild-mode 3 --tuning 0,$b,-1; done
ncu --set full --clock-control none --import-source on -k regex:"k_gather_scs" --launch-skip 1 -c 1 -o gpurun_out/${tag}_gather python tools/bench_phases.py --configs c2 --steps 2 --shuffling 0 --rebuild-mode 3 --tuning 0,7,-1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv python tools/bench_phases.py --configs c2 --steps 3 --shuffling 0 --rebuild-mode 3 --tuning 0,7,-1 > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.txt 2>/dev/null; head -8 gpurun_out/${tag}_launches.txt
