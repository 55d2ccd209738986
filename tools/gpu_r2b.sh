#!/bin/bash
# rebuild A/B: stage vs single-pass gather (chunk order on/off)
tag=${1:-r2b}
mkdir -p gpurun_out
python -m pytest tests/test_structures_gpu.py tests/test_zz_mirror_gpu.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -5 gpurun_out/${tag}_pytest.log
for args in "--rebuild-mode 1" "--rebuild-mode 2 --chunk-order 1" "--rebuild-mode 2 --chunk-order 0" "--rebuild-mode 2 --chunk-order 1 --shuffling 0" "--rebuild-mode 1 --shuffling 0"; do
  python tools/bench_phases.py --configs c2 --steps 12 $args 2>>gpurun_out/${tag}_phases.err | tee -a gpurun_out/${tag}_phases.jsonl | cut -c1-420
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv python tools/bench_phases.py --configs c2 --steps 2 --shuffling 0 > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.txt 2>/dev/null; head -40 gpurun_out/${tag}_launches.txt
