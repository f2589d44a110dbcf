"""Aggregates an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: launches, total time, share.
usage: python scripts/launch_share.py gpurun_out/launches.csv [skip_launches] > profiles/<round>_launch_shares.md

The per-launch times of such a pass are cold-cache and serialised, so only the SHARE of each kernel is meaningful; the
absolute step time comes from bench.py's CUDA events."""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name if len(name) < 110 else name[:107] + "..."


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    i_name, i_metric, i_val, i_id = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
    for r in rd:
        if r[i_metric] == "gpu__time_duration.sum" and int(r[i_id]) >= skip:
            rows.append((short(r[i_name]), float(r[i_val].replace(",", ""))))
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"source: `{path}` -- {len(rows)} launches, {total / 1e6:.2f} ms summed (serialised, cold cache)\n")
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% |")


if __name__ == "__main__":
    main()
