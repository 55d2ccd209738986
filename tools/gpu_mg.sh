# multi-GPU checks on an N-GPU box: parity worker at 1/2/4 ranks, full PIC step and headline bench scaling
N=$(nvidia-smi -L | wc -l)
tag=${1:-r1h}
export MGPU_UNVERIFIED=1   # also run the scenarios that have not passed on hardware yet (tests/mgpu_worker.py)
mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py -x -q > gpurun_out/${tag}_multigpu.log 2>&1; tail -15 gpurun_out/${tag}_multigpu.log
for n in 1 2 4 8; do
 [ $n -le $N ] || continue
 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/bench_picstep.py --steps 10 2>gpurun_out/${tag}_picstep_$n.err | tail -1 | tee -a gpurun_out/${tag}_picstep.jsonl
done
for n in 2 4 8; do
 [ $n -le $N ] || continue
 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee -a gpurun_out/${tag}_bench_scale.jsonl | cut -c1-330
done
