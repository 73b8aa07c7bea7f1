# Round 2, second 2-GPU call (gpurun --gpus 2 --timeout 1500 -- 'bash tools/r02_call6.sh'): the multi-rank parity check on the
# current library (two-phase merged rows in the fused kernel), bench at N=2 with the NVLink counters, and the rehearsal of
# the distributed setup (what 512^3 on 8 GPUs uses) at 256^3: must reproduce the one-GPU run's level sizes and 9 iterations.
mkdir -p gpurun_out
set -x
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 tests/multigpu_check.py > gpurun_out/r02c_mg2.log 2>&1; echo "multigpu_check exit $?"
grep -E "MULTIGPU_OK|FAILED|bounded wait|Error|error" gpurun_out/r02c_mg2.log | head -12; tail -3 gpurun_out/r02c_mg2.log | cut -c1-300
SAENA_BENCH_NVLINK=1 SAENA_BENCH_AB=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --no-cpu-baseline 2> gpurun_out/r02c_bench_n2.err | tee gpurun_out/r02c_bench_n2.json | cut -c1-300
echo "bench exit $?"; grep -E "rank|Error|error|FAILED|fallback" gpurun_out/r02c_bench_n2.err | tail -8
python - <<'P'
import json
for l in open("gpurun_out/r02c_bench_n2.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['iterations'], d.get('halo_fallback')); print({k:v for k,v in d['vcycle_graph'].items() if k!='halo_autotune'}); print([round(x,3) for x in d['vcycle_levels']['level_share']]); print(d.get('halo_overlap'))
P
SAENA_B200_MERGED_SPLIT=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --no-cpu-baseline 2> gpurun_out/r02c_bench_n2_nosplit.err | tee gpurun_out/r02c_bench_n2_nosplit.json | cut -c1-200
python - <<'P'
import json
for l in open("gpurun_out/r02c_bench_n2_nosplit.json"):
    if l.startswith('{'):
        d=json.loads(l); print("no merged split:", d['n_gpus'], d['ms_per_step'], d['iterations']); print([round(x,3) for x in d['vcycle_levels']['level_share']])
P
SAENA_BENCH_VERBOSE=1 SAENA_BENCH_VERIFY=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --size 256 --steps 5 --no-cpu-baseline --dist-setup on 2> gpurun_out/r02c_bench_256_dist_n2.err | tee gpurun_out/r02c_bench_256_dist_n2.json | cut -c1-300
echo "dist-setup bench exit $?"; grep -E "level |setup|Error|error" gpurun_out/r02c_bench_256_dist_n2.err | tail -30
python - <<'P'
import json
for l in open("gpurun_out/r02c_bench_256_dist_n2.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['iterations'], d['rel_residual'], d['true_rel_residual'], d.get('verify')); print([(e['level'], e['rows'], e['nnz']) for e in d['levels']])
P
