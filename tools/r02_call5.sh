# Round 2, second one-GPU call: gpurun --timeout 1800 -- 'bash tools/r02_call5.sh'
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/r02b_pytest.log 2>&1; tail -3 gpurun_out/r02b_pytest.log; grep -E "^n=" gpurun_out/r02b_pytest.log | cut -c1-200
SAENA_BENCH_VERBOSE=1 timeout 600 python bench.py 2> gpurun_out/r02b_bench.err | tee gpurun_out/r02b_bench.json | cut -c1-300
python - <<'P'
import json
for l in open("gpurun_out/r02b_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("ms/solve", d["ms_per_step"], "iters", d["iterations"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("pageable_ms_per_step"))
        print("roofline", d["roofline"]["frac"], "cpu_baseline", {k: v for k, v in d.get("cpu_baseline", {}).items() if k != "sample"})
        print("mappings changed", d.get("row_mappings_changed_by_setup_autotune"))
        for e in d["levels"][:6]: print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in e.items()})
P
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/r02b_reference_arm.json 2> gpurun_out/r02b_reference_arm.err; cut -c1-1500 gpurun_out/r02b_reference_arm.json; tail -4 gpurun_out/r02b_reference_arm.err
# ncu --set full of every kernel of one PCG iteration (2 V-cycles, all levels), eager launches, only between
# cudaProfilerStart/Stop; raw page as CSV (the report itself stays on the box)
python tools/profile_kernels.py 256 > gpurun_out/r02b_profile_kernels_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled -c 260 -o /tmp/r02b_allk \
    python tools/profile_kernels.py 256 > gpurun_out/r02b_ncu_allk.log 2>&1
tail -2 gpurun_out/r02b_ncu_allk.log | cut -c1-300
ncu -i /tmp/r02b_allk.ncu-rep --page raw --csv > /tmp/r02b_allk_raw.csv 2>/dev/null
python - <<'P'
import csv
rows = list(csv.reader(open("/tmp/r02b_allk_raw.csv")))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name", "Block Size", "Grid Size") or h.split(".")[0] in (
    "gpu__time_duration", "dram__bytes_read", "dram__bytes_write", "dram__throughput", "sm__warps_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "sm__maximum_warps_per_active_cycle_pct", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "sm__throughput", "gpu__compute_memory_throughput",
    "smsp__cycles_active", "launch__shared_mem_per_block_static", "lts__t_bytes", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active", "smsp__inst_executed", "sm__inst_executed_pipe_fp64")]
with open("gpurun_out/r02b_allk_raw.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep])
print(len(rows) - 2, "launches,", len(keep), "columns kept")
P
timeout 600 python tools/setup_time.py 256 > gpurun_out/r02b_setup_time.json 2> gpurun_out/r02b_setup_time.err; cat gpurun_out/r02b_setup_time.json; tail -3 gpurun_out/r02b_setup_time.err
timeout 400 python tools/fused_restrict_bench.py 256 > gpurun_out/r02b_fused_restrict.jsonl 2> gpurun_out/r02b_fused_restrict.err; cat gpurun_out/r02b_fused_restrict.jsonl
timeout 400 python tools/profile_fused.py --n 256 --ranks 8 --rank 3 > gpurun_out/r02b_profile_fused_n8r3.jsonl 2> gpurun_out/r02b_profile_fused.err; cat gpurun_out/r02b_profile_fused_n8r3.jsonl | cut -c1-250
timeout 400 python bench.py --workload unstructured2d --steps 5 --no-cpu-baseline 2> gpurun_out/r02b_bench_unstructured.err | tee gpurun_out/r02b_bench_unstructured.json | cut -c1-300
