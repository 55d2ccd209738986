#!/bin/bash
# A/B timing of library variants: tools/gpu_ab.sh name1 name2 ... ("base" = the in-tree library)
for v in "$@"; do
  if [ "$v" = base ]; then unset PUMIPIC_B200_LIB; else export PUMIPIC_B200_LIB=$PWD/pumi-pic_b200/_variants/lib_$v.so; fi
  python bench.py --no-cpu-baseline --no-e2e $BENCH_ARGS 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e9,2), 'G/s', round(d['ms_per_step'],4), 'ms frac', round(d['roofline']['frac'],3))"
done
