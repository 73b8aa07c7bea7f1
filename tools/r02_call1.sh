# The first one-GPU call of the next session, everything NEXT_STEPS.md lists for one GPU, each step bounded:
#   gpurun --timeout 2400 -- 'bash tools/r02_call1.sh'
# Outputs under gpurun_out/: r02_pytest.log, r02_bench.json/.err, r02_launches.csv, r02_transfers*, sanitize_*.log
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest.log 2>&1; tail -5 gpurun_out/r02_pytest.log
SAENA_BENCH_VERBOSE=1 SAENA_BENCH_AUTOTUNE_MAP=1 timeout 600 python bench.py --steps 5 2> gpurun_out/r02_bench.err | tee gpurun_out/r02_bench.json | cut -c1-400
python - <<'P'
import json
for l in open("gpurun_out/r02_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("ms/solve", d["ms_per_step"], "iters", d["iterations"], "rel", d["rel_residual"], "true", d.get("true_rel_residual"))
        print("roofline", d["roofline"]["frac"], "cpu_baseline", d.get("cpu_baseline", {}).get("value"))
        print("mapping_autotune", json.dumps(d.get("mapping_autotune"))[:1500])
P
# configs[4]'s shape on one GPU with the same in-run A/B (irregular rows: where mapping 101 should pay)
SAENA_BENCH_AUTOTUNE_MAP=1 timeout 600 python bench.py --workload unstructured2d --steps 5 --no-cpu-baseline 2> gpurun_out/r02_bench_unstructured.err \
  | tee gpurun_out/r02_bench_unstructured.json | cut -c1-300
python - <<'P'
import json
for l in open("gpurun_out/r02_bench_unstructured.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("unstructured ms/solve", d["ms_per_step"], "mapping_autotune", json.dumps(d.get("mapping_autotune"))[:1500])
P
# ncu launch list of one bench solve (only after the plain run above exited 0), then the transfer kernels
K='regex:spmv_|halo_pack|dot_kernel|pcg_|cheb_first|negate_copy|coarsest_kernel|carry_scalar|cg_p_|scale_vector|widen_ghost'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 2380 -c 800 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --no-cpu-baseline > gpurun_out/r02_ncu1.log 2>&1
tail -2 gpurun_out/r02_ncu1.log | cut -c1-300; python tools/summarize_launches.py gpurun_out/r02_launches.csv 2>/dev/null | head -30
timeout 600 bash tools/profile_transfers.sh > gpurun_out/r02_profile_transfers.log 2>&1; tail -12 gpurun_out/r02_profile_transfers.log
timeout 900 bash tools/sanitize.sh > gpurun_out/r02_sanitize.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitize.log gpurun_out/sanitize_*.log | head -12
