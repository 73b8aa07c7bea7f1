# usage: bash tools/mg_bench.sh N   (under gpurun --gpus N): parity check, then bench with the in-run A/Bs
N=${1:-4}
set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 tests/multigpu_check.py > gpurun_out/mg${N}g.log 2>&1; echo "check exit $?"
grep -E "MULTIGPU_OK|FAILED|Error|assert" gpurun_out/mg${N}g.log | head -20
bash tools/bench_n.sh $N
