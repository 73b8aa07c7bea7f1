# Round 2, the 4-GPU call (gpurun --gpus 4 --timeout 1500 -- 'bash tools/r02_call3.sh'): what hung in round 1's driver run
# (bench.py --gpus 4 at HEAD), the multi-rank parity check and the reference's own 4-rank layout on 4 GPUs, the
# drop-in on 4 MPI ranks, and two A/Bs of layout knobs.
mkdir -p gpurun_out
set -x
nvidia-smi -L | head -4
SAENA_BENCH_AB=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 10 --no-cpu-baseline 2> gpurun_out/r02_bench_n4.err | tee gpurun_out/r02_bench_n4.json | cut -c1-300
echo "bench exit $?"; grep -E "rank|Error|error|FAILED|fallback" gpurun_out/r02_bench_n4.err | tail -8
python - <<'P'
import json
for l in open("gpurun_out/r02_bench_n4.json"):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['iterations'], d.get('halo_fallback')); print({k:v for k,v in d['vcycle_graph'].items() if k!='halo_autotune'}); print([round(x,3) for x in d['vcycle_levels']['level_share']]); print(d.get('row_mappings_changed_by_setup_autotune'))
P
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556 tests/multigpu_check.py > gpurun_out/r02_mg4.log 2>&1; echo "multigpu_check exit $?"
grep -E "MULTIGPU_OK|FAILED|bounded wait|Error|error" gpurun_out/r02_mg4.log | head -12; tail -3 gpurun_out/r02_mg4.log | cut -c1-300
timeout 600 python -m pytest tests/test_multigpu.py tests/test_public_api_dropin.py -q -m gpu -k "4 and not distributed_path" > gpurun_out/r02_pytest_4gpu.log 2>&1; tail -12 gpurun_out/r02_pytest_4gpu.log | cut -c1-400
for knob in "SAENA_B200_MERGED_SPLIT=1" "SAENA_B200_MERGE_ABOVE=0.08"; do
  env $knob timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 10 --no-cpu-baseline 2> gpurun_out/r02_bench_n4_knob.err | tee gpurun_out/r02_bench_n4_$knob.json | cut -c1-200
  python - "$knob" <<'P'
import json, sys
for l in open(f"gpurun_out/r02_bench_n4_{sys.argv[1]}.json"):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d['n_gpus'], d['ms_per_step'], d['iterations']); print([round(x,3) for x in d['vcycle_levels']['level_share']])
P
done
