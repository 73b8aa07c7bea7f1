# Round 2, the 8-GPU call (gpurun --gpus 8 --timeout 1500 -- 'bash tools/r02_call4.sh'):
#   1. BASELINE.json configs[2]: 3D Poisson 512^3 = 134 217 728 unknowns row-partitioned over 8 B200, hierarchy built by the
#      distributed setup (saena_b200/sa_setup_dist.py), AMG-PCG to 1e-8 -- the north_star's target configuration
#   2. configs[1] (256^3) on 8 GPUs: the strong-scaling point of the driver's SCALE run, with the agglomeration threshold sweep
mkdir -p gpurun_out
set -x
nvidia-smi -L | wc -l
SAENA_BENCH_VERBOSE=1 SAENA_BENCH_VERIFY=1 timeout 1100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --size 512 --steps 5 --no-cpu-baseline 2> gpurun_out/r02_bench_512_n8.err | tee gpurun_out/r02_bench_512_n8.json | cut -c1-400
echo "512^3 bench exit $?"
grep -E "level |setup|Error|error|rank|fallback" gpurun_out/r02_bench_512_n8.err | tail -40 | cut -c1-200
python - <<'P'
import json
for l in open("gpurun_out/r02_bench_512_n8.json"):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['n_gpus'], 'ms/solve', d['ms_per_step'], 'iters', d['iterations'], 'rel', d['rel_residual'], 'true', d['true_rel_residual'], 'value', d['value'], 'e2e', d['e2e'].get('ms_per_step'))
        print('fallback', d.get('halo_fallback'), 'verify', (d.get('verify') or {}).get('worst'), (d.get('verify') or {}).get('ok'))
        print('halo', d.get('halo_overlap'))
        print('shares', [round(x,3) for x in d['vcycle_levels']['level_share']])
        print('roofline', d['roofline'])
        for e in d['levels']: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in e.items()})
P
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 10 --no-cpu-baseline 2> gpurun_out/r02_bench_n8.err | tee gpurun_out/r02_bench_n8.json | cut -c1-300
echo "256^3 N=8 bench exit $?"; grep -E "rank|Error|error|FAILED|fallback" gpurun_out/r02_bench_n8.err | tail -8
python - <<'P'
import json
for l in open("gpurun_out/r02_bench_n8.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['iterations'], d.get('halo_fallback'), d.get('agglomerate_sweep')); print([round(x,3) for x in d['vcycle_levels']['level_share']]); print(d.get('halo_overlap')); print(d.get('row_mappings_changed_by_setup_autotune'))
P
