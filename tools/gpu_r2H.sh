#!/bin/bash
# final 1-GPU validation of the round: full GPU test suite, smoke, the bench lines, ncu of the shipped
# headline kernel and of the rebuild's gather, launch lists of the bench and of the full PIC step
tag=${1:-r2H}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2>gpurun_out/${tag}_bench_reference.err; tail -c 600 gpurun_out/${tag}_bench_reference_arm.json; echo
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2>gpurun_out/${tag}_bench.err; python -c "
import json; r=json.loads(open('gpurun_out/${tag}_bench_n1.json').read().strip().splitlines()[-1])
print('value',r['value'],'ms',r['ms_per_step'],'frac',r['roofline']['frac'],'e2e',r['e2e']['value'],'cpu',r['cpu_baseline']['value'],'parity',r['parity']); p=r['picstep']; print('picstep',p['ms_per_step'],p['phase_ms'],p.get('parity'))"
ncu --set full --clock-control none --import-source on -k regex:"k_walk_scs" --launch-skip 6 -c 1 -o gpurun_out/${tag}_walk python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-picstep --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gather_scs|k_invmap|k_hist_kept" --launch-skip 9 -c 3 -o gpurun_out/${tag}_rebuild python tools/bench_picstep.py --steps 3 --warmup 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-picstep --no-graph > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${tag}_launches_bench.csv 8 > gpurun_out/${tag}_launches_bench.txt 2>/dev/null; cat gpurun_out/${tag}_launches_bench.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches_picstep.csv python tools/bench_picstep.py --steps 3 --warmup 1 > /dev/null 2>&1
python tools/launch_list.py gpurun_out/${tag}_launches_picstep.csv > gpurun_out/${tag}_launches_picstep.txt 2>/dev/null; head -40 gpurun_out/${tag}_launches_picstep.txt
timeout 300 python tools/bench_phases.py --configs c2 --steps 12 2>/dev/null | tee gpurun_out/${tag}_phases_c2.jsonl | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['config'],{k:round(v['median_ms'],4) for k,v in r['phases'].items()})"
ls -la gpurun_out/${tag}_*.ncu-rep
