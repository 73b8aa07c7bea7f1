# Round 2: two GPUs, the library at HEAD (one fence per pack CTA, mapping walk over the row-group mappings, merged layout from
# 8 % of rows with remote entries): parity + fault injection, then the bench line.
mkdir -p gpurun_out
set -x
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 tests/multigpu_check.py > gpurun_out/r02f_mg2.log 2>&1; echo "multigpu_check exit $?"
grep -E "MULTIGPU_OK|FAILED|bounded wait|Error|error" gpurun_out/r02f_mg2.log | head -12; tail -3 gpurun_out/r02f_mg2.log | cut -c1-300
SAENA_BENCH_AB=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --no-cpu-baseline 2> gpurun_out/r02f_bench_n2.err | tee gpurun_out/r02f_bench_n2.json | cut -c1-300
echo "bench exit $?"; grep -E "rank|Error|error|FAILED|fallback" gpurun_out/r02f_bench_n2.err | tail -8
python - <<'P'
import json
for l in open("gpurun_out/r02f_bench_n2.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['iterations'], d.get('halo_fallback')); print({k:v for k,v in d['vcycle_graph'].items() if k!='halo_autotune'}); print([round(x,3) for x in d['vcycle_levels']['level_share']]); print(d.get('halo_overlap')); print(d.get('row_mappings_changed_by_setup_autotune'))
P
