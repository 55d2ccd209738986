#!/bin/bash
# balancer selection with block-level quota reservation: tests + timing; the final bench lines of the round
tag=${1:-r2M}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_lb_gpu.py -m gpu -x -q > gpurun_out/${tag}_lb_tests.log 2>&1; tail -3 gpurun_out/${tag}_lb_tests.log | cut -c1-300
timeout 600 python -m pytest tests/test_multigpu.py -x -q -k "one_gpu" > gpurun_out/${tag}_shared_gpu_tests.log 2>&1; tail -3 gpurun_out/${tag}_shared_gpu_tests.log | cut -c1-600
timeout 300 python tools/bench_lb.py 2>gpurun_out/${tag}_lb.err | tee gpurun_out/${tag}_balancer_kernels.json; tail -2 gpurun_out/${tag}_lb.err
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2>gpurun_out/${tag}_bench.err; python -c "
import json; r=json.loads(open('gpurun_out/${tag}_bench_n1.json').read().strip().splitlines()[-1])
print('value',r['value'],'ms',r['ms_per_step'],'frac',r['roofline']['frac'],'e2e',r['e2e']['value'],'cpu',r['cpu_baseline']['value'],'parity',r['parity']['mismatch']); p=r['picstep']; print('picstep',p['ms_per_step'],p.get('parity',{}).get('mismatch'),p.get('full_size_check'))"
tail -2 gpurun_out/${tag}_bench.err
