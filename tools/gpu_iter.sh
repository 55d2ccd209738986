#!/bin/bash
# One GPU iteration on the walk kernel: parity tests, a short bench, an ncu capture named $1.
tag=${1:-iter}
python -m pytest tests/test_walk_kernels_gpu.py -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value']/1e9, 'G/s', d['ms_per_step'], 'ms frac', d['roofline']['frac'], d['detail'])"
ncu --set full --clock-control none --import-source on -k regex:k_walk_scs -s 3 -c 1 -o gpurun_out/$tag python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
