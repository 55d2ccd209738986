#!/bin/bash
tag=${1:-r2t}
mkdir -p gpurun_out
for dbg in 0 1 2; do
 PUMIPIC_PACK_DEBUG=$dbg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$dbg tools/bench_picstep.py --steps 12 --timing 2>gpurun_out/${tag}_picstep_$dbg.err | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); t=r.get('library_phase_avg_ms_rank0'); print('debug $dbg: pack kernel', t.get('migration pack kernel'), 'pack+stores', t.get('migration pack + peer stores'), 'ms/step', round(r['ms_per_step'],4))"
done
