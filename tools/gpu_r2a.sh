#!/bin/bash
# round-2 first call: state of the tree on hardware
tag=${1:-r2a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${tag}_pytest_gpu.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench.json
python tools/bench_phases.py --configs c2 > gpurun_out/${tag}_phases_c2.json 2> gpurun_out/${tag}_phases.err; cut -c1-1500 gpurun_out/${tag}_phases_c2.json
python tools/bench_lb.py > gpurun_out/${tag}_lb.json 2> gpurun_out/${tag}_lb.err; cat gpurun_out/${tag}_lb.json
timeout 300 python tools/bench_c3_sweep.py --quick --iters 10 > gpurun_out/${tag}_c3_sweep.jsonl 2> gpurun_out/${tag}_c3_sweep.err; cut -c1-260 gpurun_out/${tag}_c3_sweep.jsonl
