# Round 2, first one-GPU call, each step bounded:
#   gpurun --timeout 2700 -- 'bash tools/r02_call1.sh'
# Outputs under gpurun_out/: r02_pytest.log, r02_bench.json/.err, r02_launches.csv, r02_allk_raw.csv, r02_transfers*, sanitize_*.log
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1700 python -m pytest tests -x -q -m gpu -s > gpurun_out/r02_pytest.log 2>&1; tail -5 gpurun_out/r02_pytest.log; grep -E "^n=|passed|failed" gpurun_out/r02_pytest.log | tail -5
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
# ncu launch list of one bench solve (only after the plain run above exited 0)
K='regex:spmv_|halo_pack|dot_kernel|pcg_|cheb_first|negate_copy|coarsest_kernel|carry_scalar|cg_p_|scale_vector|widen_ghost'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 2380 -c 800 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --no-cpu-baseline > gpurun_out/r02_ncu1.log 2>&1
tail -2 gpurun_out/r02_ncu1.log | cut -c1-300; python tools/summarize_launches.py gpurun_out/r02_launches.csv 2>/dev/null | head -30
# ncu --set full of EVERY kernel of one V-cycle + one PCG iteration (the first solve of the driver), raw page as CSV
python tools/profile_kernels.py 256 > gpurun_out/profile_kernels_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -c 125 -o /tmp/r02_allk \
    python tools/profile_kernels.py 256 > gpurun_out/r02_ncu_allk.log 2>&1
tail -2 gpurun_out/r02_ncu_allk.log | cut -c1-300
ncu -i /tmp/r02_allk.ncu-rep --page raw --csv > gpurun_out/r02_allk_raw.csv 2>/dev/null; ls -la /tmp/r02_allk.ncu-rep gpurun_out/r02_allk_raw.csv
timeout 600 bash tools/profile_transfers.sh > gpurun_out/r02_profile_transfers.log 2>&1; tail -12 gpurun_out/r02_profile_transfers.log
timeout 700 bash tools/sanitize.sh > gpurun_out/r02_sanitize.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitize.log gpurun_out/sanitize_*.log | head -12
# configs[4]'s shape on one GPU with the same in-run A/B (irregular rows: where mapping 101 should pay)
SAENA_BENCH_AUTOTUNE_MAP=1 timeout 500 python bench.py --workload unstructured2d --steps 5 --no-cpu-baseline 2> gpurun_out/r02_bench_unstructured.err \
  | tee gpurun_out/r02_bench_unstructured.json | cut -c1-300
python - <<'P'
import json
for l in open("gpurun_out/r02_bench_unstructured.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("unstructured ms/solve", d["ms_per_step"], "mapping_autotune", json.dumps(d.get("mapping_autotune"))[:1500])
P
