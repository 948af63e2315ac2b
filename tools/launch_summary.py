"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel name the
number of launches, total and mean device time and the share of the captured window.
usage: launch_summary.py launches.csv [> profiles/xxx.md]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = []
    with open(sys.argv[1]) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] == "us":
            v *= 1e3
        elif r[ui] == "ms":
            v *= 1e6
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"<.*", "", name).replace("void ", "").strip()
        rows.append((name, v))
    tot = sum(v for _, v in rows)
    agg = OrderedDict()
    for n, v in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    print("| kernel | launches | total us | mean us | share |")
    print("|---|---|---|---|---|")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.2f | %.1f %% |" % (n, c, v / 1e3, v / c / 1e3, 100 * v / tot))
    print("| **all** | %d | %.1f | | 100 %% |" % (len(rows), tot / 1e3))


if __name__ == "__main__":
    main()
