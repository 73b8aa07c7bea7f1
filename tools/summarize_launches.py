"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel summary
(markdown).  usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/x.md"""
import collections
import csv
import io
import re
import sys


def main(path):
    text = open(path).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        name = r["Kernel Name"]
        m = re.match(r"(?:void )?(\w+)(<[^>]*>)?", name)
        short = (m.group(1) + (m.group(2) or "")) if m else name[:40]
        key = (short, r["Grid Size"])
        t = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    print(f"launches: {len(rows)}, summed device time {tot / 1e6:.3f} ms (cold-cache, serialised: compare shares)\n")
    print("| kernel | grid | launches | total ms | share | avg us |")
    print("|---|---|---:|---:|---:|---:|")
    for (k, g), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {g} | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f}% | {t / c / 1e3:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
