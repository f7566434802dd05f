"""Counts of the SASS instructions that prove what the kernels are made of, per kernel of the built
library -> profiles/r02_sass_summary.md (cuobjdump works without a GPU):
    python tools/sass_summary.py
UTMALDG = TMA tensor load, SYNCS = mbarrier, LDGSTS = cp.async, IDP.2A = dp2a masked sums,
REDUX = warp reductions, ATOMS = shared-memory atomics (the per-warp median histograms),
STG.E.128 / LDG.E.128 = 128-bit global accesses, DFMA = the one-DFMA flat-field fast path."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "magnify_b200", "libmagnify_b200.so")
WHAT = ["UTMALDG", "SYNCS", "LDGSTS", "IDP.2A", "REDUX", "ATOMS", "STG.E.128", "LDG.E.128", "LDS.128", "DFMA", "DADD", "MUFU.RCP64H"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    arch = re.findall(r"arch = (sm_\w+)", out)
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("mgb::", "").replace("void ", "")
            kernels[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[name]["total"] += 1
            for w in WHAT:
                if op == w or op.startswith(w + ".") or (w in ("STG.E.128", "LDG.E.128", "LDS.128") and w in op):
                    kernels[name][w] += 1
    lines = ["# SASS instruction counts per kernel (round 2)", "",
             f"`cuobjdump -sass magnify_b200/libmagnify_b200.so`, architectures in the file: {sorted(set(arch))}.",
             "Counts are static instructions in the kernel body (one row per template instantiation).", "",
             "| kernel | total | " + " | ".join(WHAT) + " |", "|---|---|" + "---|" * len(WHAT)]
    totals = collections.Counter()
    for k, c in kernels.items():
        lines.append(f"| `{k}` | {c['total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WHAT) + " |")
        totals.update(c)
    lines.append("| **all kernels** | %d | " % totals["total"] + " | ".join(str(totals[w]) for w in WHAT) + " |")
    path = os.path.join(ROOT, "profiles", "r02_sass_summary.md")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(path, {w: totals[w] for w in WHAT})


if __name__ == "__main__":
    sys.exit(main())
