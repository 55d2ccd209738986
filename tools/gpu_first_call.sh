#!/bin/bash
# First GPU call of a round: everything that was written without GPU time, logs into gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh r2a'
tag=${1:-r2a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
# the quarantined tests for real: --runxfail turns their xfail markers off
python -m pytest tests/test_zz_lb_gpu.py tests/test_zz_mirror_gpu.py -m gpu -q --runxfail > gpurun_out/${tag}_quarantine.log 2>&1; tail -15 gpurun_out/${tag}_quarantine.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench.json
python tools/bench_lb.py > gpurun_out/${tag}_lb.json 2> gpurun_out/${tag}_lb.err; cat gpurun_out/${tag}_lb.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_bench_reference.json
python tools/bench_c3_sweep.py --quick --iters 10 > gpurun_out/${tag}_c3_sweep.jsonl 2> gpurun_out/${tag}_c3_sweep.err; cut -c1-260 gpurun_out/${tag}_c3_sweep.jsonl
