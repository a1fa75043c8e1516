"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, mean time, share of K1+K2+K3.

    python tools/ncu_launch_summary.py profiles/<name>.csv "<command that was profiled>" > profiles/<name>_summary.txt
"""
import csv
import sys
from collections import defaultdict


def main(path: str, what: str) -> None:
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    t = defaultdict(list)
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        t[r[ki]].append(v / 1e3 if r[ui] in ("ns", "nsecond") else v)
    hot = {k: v for k, v in t.items() if any(s in k for s in ("k1_lse", "k2_lattice", "k3_grad"))}
    tot = sum(sum(v) / len(v) for v in hot.values())
    print(f"ncu launch list, `{what}`, per-launch gpu__time_duration (cold cache, serialised):\n")
    for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
        share = f"{100 * (sum(v) / len(v)) / tot:5.1f}%" if k in hot else "    -"
        print(f"{k[:70]:70s} n={len(v):3d} mean {sum(v) / len(v):9.1f} us  share-of-K1+K2+K3 {share}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?")
