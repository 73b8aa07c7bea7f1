"""bench.py JSON lines (one file per N) -> a markdown report: whole-job throughput, ms per solve, speed-up and
parallel efficiency against the smallest N given, e2e, iterations / residuals, clocks, halo overlap, the in-run A/Bs
and the per-level V-cycle cost.

    python tools/scaling_report.py gpurun_out/bench256_final2.json gpurun_out/bench_n2g.json ... > profiles/rNN_scaling.md
"""
import json
import sys


def load(path):
    for line in open(path):
        line = line.strip()
        if line.startswith("{"):
            d = json.loads(line)
            if "ms_per_step" in d:
                return d
    raise SystemExit(f"{path}: no bench line")


def main():
    runs = sorted((load(p) for p in sys.argv[1:]), key=lambda d: d["n_gpus"])
    if not runs:
        raise SystemExit(__doc__)
    base = runs[0]
    print(f"# {base['metric']}\n")
    print("| GPUs | workload | ms / solve | " + base["unit"] + " | speed-up | efficiency | e2e " + base["unit"] +
          " | iterations | rel. residual (recurrence / recomputed) | SM MHz (reasons) |")
    print("|---:|---|---:|---:|---:|---:|---:|---:|---|---|")
    for d in runs:
        n = d["n_gpus"]
        same = d["config"]["workload"] == base["config"]["workload"]
        sp = d["value"] / base["value"]
        eff = sp / (n / base["n_gpus"])
        clk = d.get("clocks", {})
        print(f"| {n} | {d['config']['workload'].split(',')[0]} | {d['ms_per_step']:.2f} | {d['value']:.1f} | "
              f"{sp:.2f}{'' if same else ' (other workload)'} | {eff:.2f} | {d['e2e']['value']:.1f} | {d.get('iterations')} | "
              f"{d.get('rel_residual', float('nan')):.2e} / "
              f"{(d.get('true_rel_residual') if d.get('true_rel_residual') is not None else float('nan')):.2e} | "
              f"{clk.get('sm_mhz')} ({', '.join(clk.get('reasons', [])) or '-'}) |")
    print()
    for d in runs:
        n = d["n_gpus"]
        extra = []
        h = d.get("halo_overlap")
        if h:
            extra.append(f"halo at level 0: full {h['spmv_full_ms']:.4f} ms, compute only {h['spmv_local_only_ms']:.4f}, "
                         f"exchange only {h['pack_exchange_only_ms']:.4f} -> hidden {h['hidden_frac']:.2f}; "
                         f"{h['ghost_values_per_rank']} ghost values per rank ({h['ghost_dtype']})")
        g = d.get("vcycle_graph", {})
        ab = [f"{k.replace('_ms_per_step', '')} {v:.2f} ms" for k, v in g.items() if k.endswith("ms_per_step")]
        if ab:
            extra.append("same solves, other launch paths: " + ", ".join(ab))
        m = d.get("mapping_autotune")
        if m:
            extra.append(f"row-mapping autotune: {m['ms_per_step_heuristic']:.2f} -> {m['ms_per_step_autotuned']:.2f} ms "
                         f"({len(m['changed'])} operators changed)")
        v = d.get("verify")
        if v:
            extra.append(f"hierarchy properties (symmetry, R = P^T): worst {v.get('worst', float('nan')):.1e}, ok = {v.get('ok')}")
        r = d.get("roofline", {})
        if r:
            extra.append(f"dominant kernel ({r.get('kernel')}): {r['achieved']:.0f} GB/s = {r['frac']:.2f} of {r['peak']:.0f}")
        if extra:
            print(f"**N = {n}**: " + "; ".join(extra) + "\n")
    print("## ms per V-cycle spent on each level (entered-at-level differences, eager launches, max over ranks)\n")
    L = max(len(d["vcycle_levels"]["level_share"]) for d in runs)
    print("| level | " + " | ".join(f"N={d['n_gpus']}" for d in runs) + " |")
    print("|---:|" + "---:|" * len(runs))
    for l in range(L):
        cells = [f"{d['vcycle_levels']['level_share'][l]:.3f}" if l < len(d["vcycle_levels"]["level_share"]) else "" for d in runs]
        print(f"| {l} | " + " | ".join(cells) + " |")
    print("| total | " + " | ".join(f"{d['vcycle_levels']['entered_at_level'][0]:.2f}" for d in runs) + " |")


if __name__ == "__main__":
    main()
