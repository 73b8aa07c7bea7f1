python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 tests/multigpu_check.py > gpurun_out/mg4.log 2>&1
grep -E "MULTIGPU_OK|pcg iters|FAILED" gpurun_out/mg4.log | head
cat gpurun_out/multigpu_fail_rank0.txt 2>/dev/null | tail -30
