nvidia-smi -L | wc -l
python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -3
for n in 1 2 4; do
 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/bench_picstep.py --steps 10 2>/dev/null | tail -1 | tee -a gpurun_out/r1f_picstep.jsonl
done
for n in 2 4; do
 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 3 2>/dev/null | tail -1 | tee -a gpurun_out/r1f_bench_scale.jsonl
done
