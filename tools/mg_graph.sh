# usage: bash tools/mg_graph.sh N   (under gpurun --gpus N): parity check + bench with the graph/eager A-B
N=${1:-2}
set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 tests/multigpu_check.py > gpurun_out/mg${N}g.log 2>&1; echo "check exit $?"
grep -E "MULTIGPU_OK|FAILED|replayed|Error|assert" gpurun_out/mg${N}g.log | head -20
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | head -8
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 2> gpurun_out/bench_n${N}g.err | tee gpurun_out/bench_n${N}g.json | cut -c1-400
echo "bench exit $?"
grep -E "setup|Error|error|FAILED" gpurun_out/bench_n${N}g.err | tail -5
python - <<'P'
import json,sys
for l in open(f"gpurun_out/bench_n{sys.argv[1] if len(sys.argv)>1 else 2}g.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['vcycle_graph'], d['halo_overlap'])
P
