N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 tests/multigpu_check.py > gpurun_out/mg$N.log 2>&1
echo "exit $?"
grep -E "MULTIGPU_OK|pcg iters|FAILED" gpurun_out/mg$N.log | head -12
cat gpurun_out/multigpu_fail_rank0.txt 2>/dev/null | tail -20
