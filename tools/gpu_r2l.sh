#!/bin/bash
# profiles of the shipped headline kernel + compute-sanitizer passes over small fixtures
tag=${1:-r2l}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_walk_scs" --launch-skip 6 -c 1 -o gpurun_out/${tag}_walk python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-picstep --no-graph > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-picstep --no-graph > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${tag}_launches_bench.csv 8 > gpurun_out/${tag}_launches_bench.txt 2>/dev/null; cat gpurun_out/${tag}_launches_bench.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches_picstep.csv python tools/bench_picstep.py --steps 3 --warmup 1 > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${tag}_launches_picstep.csv > gpurun_out/${tag}_launches_picstep.txt 2>/dev/null; head -30 gpurun_out/${tag}_launches_picstep.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_structures_gpu.py tests/test_walk_kernels_gpu.py -m gpu -x -q -k "(scs_c32 and 2500) or sliced or (kuhn8 and scs) or ids_empty" > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/${tag}_memcheck.log | cut -c1-200
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_walk_kernels_gpu.py tests/test_structures_gpu.py -m gpu -x -q -k "(sliced and 1024 and kuhn8) or (kuhn3 and scs) or (gather and scs_c32 and 2500)" > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/${tag}_racecheck.log | cut -c1-200
