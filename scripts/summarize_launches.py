#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py launches.csv [header comment ...]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        a = agg.setdefault(row["Kernel Name"].split("(")[0][-70:], [0, 0.0])
        a[0] += 1
        a[1] += v
    def is_own(k):  # ncu prints the anonymous namespace as dav::<unnamed>:: or (application replay) as unnamed>::
        return "dav::" in k or "unnamed>::" in k

    own = sum(t for k, (n, t) in agg.items() if is_own(k))
    for c in sys.argv[2:]:
        print("# " + c)
    print("# own kernels (dav::) total %.1f us; shares are of that total" % own)
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        share = "%5.1f%%" % (100 * t / own) if is_own(k) else "   n/a"
        print("%-70s n=%4d total_us=%10.1f avg_us=%9.1f share=%s" % (k, n, t, t / n, share))


if __name__ == "__main__":
    main()
