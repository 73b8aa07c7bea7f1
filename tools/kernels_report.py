"""profiles/r02_allk_ncu_raw.csv (ncu --set full of every kernel of one PCG iteration, tools/profile_kernels.py) +
profiles/r02_bench_n1.json (level sizes) -> the per-kernel table: duration, DRAM traffic, algorithmic bytes
(SURVEY 8d formulas), their ratio, achieved GB/s, occupancy, registers.

    python tools/kernels_report.py > profiles/r02_kernels.md
"""
import csv
import json
import re

ROOT = "profiles/"
bench = json.loads([l for l in open(ROOT + "r02_bench_n1.json") if l.startswith("{")][0])
peak = bench["roofline"]["peak"]
levels = {e["level"]: (e["rows"], e["nnz"]) for e in bench["levels"]}
nnz_p = {}
for l in open(ROOT + "r02_fused_restrict.jsonl"):
    if l.startswith("{"):
        d = json.loads(l)
        nnz_p[d["level"]] = d["nnz_P"]
rows = list(csv.reader(open(ROOT + "r02_allk_ncu_raw.csv")))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    v, u = float(r[col[name]]), units[col[name]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return v * scale


def level_of(kernel, grid, epi):
    """which operator of which level a launch is, from its grid (rows per CTA of the mapping) and the level sizes"""
    m = re.match(r"spmv_(sell|sellp|rowgroup|vec)_kernel<(\d+)(?:, (\d+))?", kernel)
    if not m:
        return None
    kind, a = m.group(1), int(m.group(2))
    rows_per_cta = 256 if kind in ("sell", "sellp", "vec") else 256 // a
    for l, (M, nnz) in levels.items():
        if (M + rows_per_cta - 1) // rows_per_cta == grid:
            if epi in (1, 2, 3):
                return l, "A", M, nnz                      # residual / Chebyshev sweeps
            if epi == 5:
                return l, "P", M, nnz_p.get(l)
            # EPI_PLAIN: A (the Krylov loop's h = A p, level 0 only) or R of the level above
            if l == 0:
                return l, "A", M, nnz
            return l - 1, "R", M, nnz_p.get(l - 1)
    # R's rows are the coarse rows; P's rows the fine rows: both covered above (P: grid from fine M; R: grid from coarse M)
    return None


EPI_NAME = {0: "plain", 1: "residual", 2: "Chebyshev first sweep", 3: "Chebyshev sweep", 4: "Jacobi sweep", 5: "P + correction"}
seen = {}
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")
    grid = int(re.match(r"\((\d+)", r[col["Grid Size"]]).group(1))
    seen.setdefault((name, grid), []).append(r)
print("# Every kernel of one PCG iteration under `ncu --set full` — 256³ Poisson, one B200 (round 2)\n")
print("`tools/profile_kernels.py 256` (the bench hierarchy, row mappings chosen by the setup-time measurement, eager launches) under "
      "`ncu --set full --clock-control none --profile-from-start off` between cudaProfilerStart/Stop: 2 V-cycles over all 10 levels, the "
      "level-0 SpMV, the dots and the Krylov updates — 114 launches, 57 distinct (kernel, grid) pairs; raw page in "
      "`profiles/r02_allk_ncu_raw.csv` (52 of the columns). ncu replays every kernel with cold caches, so durations of the L2-sized "
      "levels are upper bounds; the bench's own CUDA-event timings are in `profiles/r02_bench_levels.md`. Algorithmic bytes = SURVEY §8d "
      f"formulas on (M, nnz); peak = {peak} GB/s (MEASURED_PEAKS.json). Registers / spills of every kernel: `profiles/r02_ptxas.md`.\n")
print("| kernel | epilogue | level, operator | grid | launches | duration µs | DRAM read + write MB | algorithmic MB | traffic ÷ algorithmic | "
      "algorithmic GB/s | ÷ peak | warps active % | regs | L1 hit % | L2 hit % |")
print("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for (name, grid), rs in seen.items():
    r = rs[0]
    t = sum(num(x, "gpu__time_duration.sum") for x in rs) / len(rs)
    rd = sum(num(x, "dram__bytes_read.sum") for x in rs) / len(rs)
    wr = sum(num(x, "dram__bytes_write.sum") for x in rs) / len(rs)
    m = re.match(r"spmv_\w+_kernel<(\d+)(?:, (\d+))?", name)
    epi = None
    if m:
        epi = int(m.group(2)) if m.group(2) is not None else int(m.group(1))
    who = level_of(name, grid, epi) if epi is not None else None
    alg = None
    if who and who[3] is not None:
        l, op, M, nnz = who
        Mfine = levels[l][0]
        if op == "A":
            alg = 12 * nnz + M * {0: 20, 1: 28, 2: 44, 3: 52}.get(epi, 20)
        elif op == "P":
            alg = 12 * nnz + Mfine * 20 + levels[l + 1][0] * 8
        else:
            alg = 12 * nnz + M * 12 + Mfine * 8
    elif name in ("dot_kernel",):
        alg = 16 * levels[0][0] if grid == 1184 else None
    elif name == "pcg_update_kernel":
        alg = 48 * levels[0][0]
    elif name == "negate_copy_kernel":
        alg = 16 * levels[0][0]
    elif name == "cheb_first_zero_kernel":
        for l, (M, _) in levels.items():
            if min((M + 1023) // 1024, 1184) == grid:
                alg, who = 32 * M, (l, "first sweep from a zero iterate", M, 0)
                break
    desc = f"L{who[0]} {who[1]}" if who else ""
    print(f"| `{name}` | {EPI_NAME.get(epi, '') if m else ''} | {desc} | {grid} | {len(rs)} | {t * 1e6:.1f} | {(rd + wr) / 1e6:.1f} | "
          f"{'' if alg is None else f'{alg / 1e6:.1f}'} | {'' if alg is None else f'{(rd + wr) / alg:.2f}'} | "
          f"{'' if alg is None else f'{alg / t / 1e9:.0f}'} | {'' if alg is None else f'{alg / t / 1e9 / peak:.2f}'} | "
          f"{float(r[col['sm__warps_active.avg.pct_of_peak_sustained_active']]):.0f} | {r[col['launch__registers_per_thread']]} | "
          f"{float(r[col['l1tex__t_sector_hit_rate.pct']]):.0f} | {float(r[col['lts__t_sector_hit_rate.pct']]):.0f} |")
print("\nReading: the three sliced-layout levels (0–2) and the row-group levels 3–4 move 0.97–1.10× their algorithmic bytes; where the ratio "
      "is below 1 the gathered vector and the epilogue streams were still in L2 from the previous launch. Levels 6–9 (≤ 1 425 rows, "
      "5–12 µs per launch, 10–20 % of the warps active) are launch-latency-bound: together < 3 % of a V-cycle.")
