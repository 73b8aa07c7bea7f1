# Round 2, last one-GPU call: gpurun --timeout 1500 -- 'bash tools/r02_call7.sh'
# the whole -m gpu suite on the final library (the device SpGEMM's tests included), the default bench line, and the bench
# with the hierarchy built through the device SpGEMM
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/r02d_pytest.log 2>&1; tail -4 gpurun_out/r02d_pytest.log | cut -c1-300; grep -E "^n=|FAILED|Error" gpurun_out/r02d_pytest.log | cut -c1-300 | head
timeout 600 python bench.py 2> gpurun_out/r02d_bench.err | tee gpurun_out/r02d_bench.json | cut -c1-300
SAENA_SETUP_SPGEMM=native timeout 600 python bench.py --steps 5 --no-cpu-baseline 2> gpurun_out/r02d_bench_native_spgemm.err | tee gpurun_out/r02d_bench_native_spgemm.json | cut -c1-300
python - <<'P'
import json
for f in ("gpurun_out/r02d_bench.json", "gpurun_out/r02d_bench_native_spgemm.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "ms/solve", d["ms_per_step"], "iters", d["iterations"], "rel", d["rel_residual"], "true", d["true_rel_residual"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("pageable_ms_per_step"))
            print("   roofline", d["roofline"]["frac"], "changed", d.get("row_mappings_changed_by_setup_autotune"))
            print("   cpu", {k: v for k, v in (d.get("cpu_baseline") or {}).items() if k != "sample"})
P
grep -E "setup\]" gpurun_out/r02d_bench.err gpurun_out/r02d_bench_native_spgemm.err
