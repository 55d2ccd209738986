#!/bin/bash
# A/B with parity: each variant runs the walk-kernel tests, then the bench
for v in "$@"; do
  if [ "$v" = base ]; then unset PUMIPIC_B200_LIB; else export PUMIPIC_B200_LIB=$PWD/pumi-pic_b200/_variants/lib_$v.so; fi
  python -m pytest tests/test_walk_kernels_gpu.py tests/test_search_gpu.py -m gpu -x -q 2>&1 | tail -1
  python bench.py --no-cpu-baseline --no-e2e --no-picstep 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],4), 'ms frac', round(d['roofline']['frac'],3))"
done
