# usage: bash tools/multigpu_run.sh N   (under gpurun --gpus N)
N=${1:-2}
set -x
nvidia-smi -L | head -8
if [ "$N" -le 4 ]; then python -m pytest tests/test_multigpu.py -x -q -k "$N" 2>&1 | tail -25; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-300
grep -E "setup|Error|error" gpurun_out/bench_n$N.err | tail -5
