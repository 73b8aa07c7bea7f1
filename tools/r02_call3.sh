# Round 2, the 4-GPU call (gpurun --gpus 4 --timeout 1200 -- 'bash tools/r02_call3.sh'): what hung in round 1's driver run
# (bench.py --gpus 4 at HEAD), the multi-rank parity check and the reference's own 4-rank layout on 4 GPUs, the
# distributed setup on 4 ranks, the drop-in on 4 MPI ranks.
mkdir -p gpurun_out
set -x
nvidia-smi -L | head -4
SAENA_BENCH_AB=1 SAENA_BENCH_VERBOSE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 10 --no-cpu-baseline 2> gpurun_out/r02_bench_n4.err | tee gpurun_out/r02_bench_n4.json | cut -c1-300
echo "bench exit $?"; grep -E "rank|Error|error|FAILED|fallback" gpurun_out/r02_bench_n4.err | tail -8
python - <<'P'
import json
for l in open("gpurun_out/r02_bench_n4.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['iterations'], d.get('halo_transport'), d.get('halo_fallback')); print(d['vcycle_graph']); print(d['vcycle_levels']); print(d.get('halo_overlap')); print(d.get('row_mappings_changed_by_setup_autotune'))
P
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556 tests/multigpu_check.py > gpurun_out/r02_mg4.log 2>&1; echo "multigpu_check exit $?"
grep -E "MULTIGPU_OK|FAILED|bounded wait|Error|error" gpurun_out/r02_mg4.log | head -12; tail -4 gpurun_out/r02_mg4.log | cut -c1-300
timeout 600 python -m pytest tests/test_multigpu.py tests/test_public_api_dropin.py -q -m gpu -k "4 and not distributed_path" > gpurun_out/r02_pytest_4gpu.log 2>&1; tail -12 gpurun_out/r02_pytest_4gpu.log | cut -c1-400
SAENA_BENCH_VERBOSE=1 SAENA_BENCH_VERIFY=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 4 --size 256 --steps 5 --no-cpu-baseline --dist-setup on 2> gpurun_out/r02_bench_256_dist_n4.err | tee gpurun_out/r02_bench_256_dist_n4.json | cut -c1-300
echo "dist-setup bench exit $?"; grep -E "level |setup|Error|error" gpurun_out/r02_bench_256_dist_n4.err | tail -30
python - <<'P'
import json
for l in open("gpurun_out/r02_bench_256_dist_n4.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['iterations'], d['rel_residual'], d['true_rel_residual'], d.get('verify'))
P
