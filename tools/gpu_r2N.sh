#!/bin/bash
tag=${1:-r2N}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_lb_gpu.py -m gpu -x -q > gpurun_out/${tag}_lb_tests.log 2>&1; tail -3 gpurun_out/${tag}_lb_tests.log | cut -c1-300
timeout 600 python -m pytest tests/test_multigpu.py -x -q -k "one_gpu" > gpurun_out/${tag}_shared_gpu_tests.log 2>&1; tail -3 gpurun_out/${tag}_shared_gpu_tests.log | cut -c1-600
timeout 300 python tools/bench_lb.py 2>gpurun_out/${tag}_lb.err | tee gpurun_out/${tag}_balancer_kernels.json; tail -2 gpurun_out/${tag}_lb.err
