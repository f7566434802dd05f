"""Sum ncu --csv metric rows per kernel name: python tools/ncu_sum.py file.csv [file2.csv ...]"""
import collections
import csv
import re
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    k_i, m_i, u_i, v_i = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows:
        if r is hdr or len(r) <= v_i or r[k_i] == "Kernel Name":
            continue
        name = re.sub(r"\(.*", "", r[k_i]).replace("void ", "").replace("mgb::", "")
        try:
            agg[name][(r[m_i], r[u_i])].append(float(r[v_i].replace(",", "")))
        except ValueError:
            pass
    print(f"## {path}")
    for name, metrics in agg.items():
        parts = []
        for (metric, unit), vals in metrics.items():
            total = sum(vals) / len(vals) if ("pct" in metric or "%" in unit) else sum(vals)
            parts.append(f"{metric}={total:.4g} {unit}")
        print(f"{name}: launches={len(next(iter(metrics.values())))}  " + "  ".join(parts))
