# BASELINE.json configs[2]: 512^3 unknowns row-partitioned over 8 B200, hierarchy built by the distributed setup.
# usage (under gpurun --gpus 8 --timeout 1500):   bash tools/bench_512.sh [N] [n] [extra bench args]
# First GPU run is still to come (written after round 1's GPU budget was spent).  Cheap rehearsal on 2 GPUs:
#   bash tools/bench_512.sh 2 256 --dist-setup on      (must give 9 iterations, as the one-GPU 256^3 run)
N=${1:-8}; n=${2:-512}; shift; shift
set -x
SAENA_BENCH_VERBOSE=1 SAENA_BENCH_NO_AB=1 SAENA_BENCH_VERIFY=1 timeout 1400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
  --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --size $n --steps 5 --no-cpu-baseline "$@" \
  2> gpurun_out/bench_${n}_n${N}.err | tee gpurun_out/bench_${n}_n${N}.json | cut -c1-400
echo "bench exit $?"
grep -E "level |aggregation|RAP|setup|Error|error" gpurun_out/bench_${n}_n${N}.err | tail -40
