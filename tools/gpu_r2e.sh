#!/bin/bash
tag=${1:-r2e}
mkdir -p gpurun_out
python -m pytest tests/test_structures_gpu.py -m gpu -x -q -k "gather or ordered" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
run() { python tools/bench_phases.py --configs c2 --steps 12 --shuffling 0 "$@" 2>>gpurun_out/${tag}_phases.err | tee -a gpurun_out/${tag}_phases.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('mode',r['rebuild_mode'],'order',r['chunk_order'],'tuning',r['tuning'],'rebuild ms',round(r['phases']['rebuild']['median_ms'],4),'min',round(r['phases']['rebuild']['min_ms'],4))"; }
for b in 1 2 3 4 6 8; do run --rebuild-mode 3 --tuning 0,$b,-1; done
run --rebuild-mode 3 --tuning 0,4,-1 --chunk-order 0
ncu --set full --clock-control none --import-source on -k regex:"k_gather_scs" --launch-skip 1 -c 1 -o gpurun_out/${tag}_gather4 python tools/bench_phases.py --configs c2 --steps 2 --shuffling 0 --rebuild-mode 3 --tuning 0,4,-1 > /dev/null 2>&1
