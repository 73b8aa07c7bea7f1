# usage: bash tools/bench_n.sh N [extra bench args]   (under gpurun --gpus N)
N=${1:-8}; shift
set -x
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --no-cpu-baseline "$@" 2> gpurun_out/bench_n${N}g.err | tee gpurun_out/bench_n${N}g.json | cut -c1-300
echo "bench exit $?"
grep -E "setup|Error|error|FAILED" gpurun_out/bench_n${N}g.err | tail -5
python - $N <<'P'
import json,sys
for l in open(f"gpurun_out/bench_n{sys.argv[1]}g.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['vcycle_graph'], d.get('agglomerate_sweep')); print(d['vcycle_levels'])
P
