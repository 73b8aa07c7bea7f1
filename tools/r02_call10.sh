# Round 2: 256^3 on N GPUs at HEAD (N = all visible), the bench line only
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --no-cpu-baseline 2> gpurun_out/r02g_bench_n$N.err | tee gpurun_out/r02g_bench_n$N.json | cut -c1-300
echo "bench exit $?"; grep -E "rank|Error|error|FAILED|fallback" gpurun_out/r02g_bench_n$N.err | tail -8
python - $N <<'P'
import json, sys
for l in open(f"gpurun_out/r02g_bench_n{sys.argv[1]}.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['iterations'], d.get('halo_fallback')); print([round(x,3) for x in d['vcycle_levels']['level_share']]); print(d.get('halo_overlap')); print(d.get('row_mappings_changed_by_setup_autotune'))
P
