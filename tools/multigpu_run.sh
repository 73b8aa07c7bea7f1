# usage: bash tools/multigpu_run.sh N   (under gpurun --gpus N)
N=${1:-2}
set -x
nvidia-smi -L | head -8
if [ "$N" -le 4 ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 tests/multigpu_check.py > gpurun_out/mg$N.log 2>&1; echo "check exit $?"
  grep -E "MULTIGPU_OK|FAILED" gpurun_out/mg$N.log | head -5
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-300
grep -E "setup|Error|error" gpurun_out/bench_n$N.err | tail -5
SAENA_B200_HALO=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --no-cpu-baseline 2> gpurun_out/bench_n${N}_nccl.err | tee gpurun_out/bench_n${N}_nccl.json | cut -c1-300
