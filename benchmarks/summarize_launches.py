"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the timed region)."""
import collections
import csv
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1)
        agg[row["Kernel Name"]][0] += 1
        agg[row["Kernel Name"]][1] += v
        tot += v
    print(f"total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare SHARES)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t / 1e6:9.3f} ms {100 * t / tot:5.1f}%  n={n:5d}  {k[:120]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
