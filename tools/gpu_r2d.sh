#!/bin/bash
tag=${1:-r2d}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_stage_pack|k_stage_unpack_scs|k_hist_kept" --launch-skip 3 -c 3 -o gpurun_out/${tag}_stage python tools/bench_phases.py --configs c2 --steps 2 --shuffling 0 --rebuild-mode 2 --chunk-order 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gather_scs|k_invmap|k_stage_pack_ordered" --launch-skip 2 -c 2 -o gpurun_out/${tag}_gather python tools/bench_phases.py --configs c2 --steps 2 --shuffling 0 --rebuild-mode 3 --tuning 0,2,-1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stage_pack_ordered" --launch-skip 1 -c 1 -o gpurun_out/${tag}_packord python tools/bench_phases.py --configs c2 --steps 2 --shuffling 0 --rebuild-mode 2 > /dev/null 2>&1
ls -la gpurun_out/
