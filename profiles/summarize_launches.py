"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (name + grid) count,
average duration and share.  usage: python profiles/summarize_launches.py <csv> [steps]"""
import collections
import csv
import sys


def main(path, steps=1):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        v = float(x["Metric Value"].replace(",", ""))
        v = {"ns": v / 1000, "us": v, "ms": v * 1000}.get(x["Metric Unit"], v / 1000)
        a = agg.setdefault(x["Kernel Name"][:64] + " grid" + x["Grid Size"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'us/step':>9} {'share':>6} {'n/step':>6} {'avg us':>8}  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1] / steps:9.1f} {100 * a[1] / tot:5.1f}% {a[0] // steps:6d} {a[1] / a[0]:8.2f}  {k}")
    print(f"{tot / steps:9.1f} total per step ({sum(a[0] for a in agg.values()) // steps} launches)")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
