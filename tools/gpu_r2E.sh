#!/bin/bash
# A/B of the chunk walk's latency variants (claim-ahead, L2 prefetch of next rows / of hop records)
tag=${1:-r2E}
mkdir -p gpurun_out
out=gpurun_out/${tag}_walk_ab.txt
: > $out
line() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],4), 'ms frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']/1e9,3) if d.get('e2e') else '')"; }
for v in base ca rows hop all base all; do
  if [ "$v" = base ]; then unset PUMIPIC_B200_LIB; else export PUMIPIC_B200_LIB=$PWD/pumi-pic_b200/_variants/lib_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-picstep 2>gpurun_out/${tag}_$v.err | tail -1 | line $v | tee -a $out
done
export PUMIPIC_B200_LIB=$PWD/pumi-pic_b200/_variants/lib_all.so
timeout 600 python -m pytest tests/test_walk_kernels_gpu.py tests/test_search_gpu.py -m gpu -x -q 2>&1 | tail -2 | tee -a $out
for v in base all; do
  if [ "$v" = base ]; then unset PUMIPIC_B200_LIB; else export PUMIPIC_B200_LIB=$PWD/pumi-pic_b200/_variants/lib_$v.so; fi
  timeout 300 python tools/bench_picstep.py --steps 20 --timing 2>>gpurun_out/${tag}_picstep.err | tail -1 | tee -a gpurun_out/${tag}_picstep_$v.jsonl | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print('$v picstep ms/step',round(r['ms_per_step'],4),{k:round(v,4) for k,v in r['phase_ms'].items()})" | tee -a $out
done
unset PUMIPIC_B200_LIB
for parts in 8 16 32; do
  timeout 300 python bench.py --no-cpu-baseline --no-picstep --e2e-parts $parts 2>>gpurun_out/${tag}_e2e.err | tail -1 | line "e2e-parts=$parts" | tee -a $out
done
