"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: summarize_launches.py launches.csv [skip_first_n_launches]
Prints: kernel (short name), launches, total us, mean us, share of the listed time."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("sj::", "")
    return name[:70]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        rows.append((short(r["Kernel Name"]), us))
    rows = rows[skip:]
    agg = OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values()) or 1.0
    print("%-72s %8s %12s %10s %7s" % ("kernel", "launches", "total us", "mean us", "share"))
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %8d %12.1f %10.1f %6.1f%%" % (k, n, us, us / n, 100 * us / total))


if __name__ == "__main__":
    main()
