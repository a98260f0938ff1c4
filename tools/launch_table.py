#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum[,..] --csv` launch list: per kernel count / avg / share."""
import collections
import csv
import sys


def main(path, last=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0] != "ID"]
    per = collections.OrderedDict()
    seq = []
    for r in rows:
        name, metric, val = r[4].split("(")[0].replace("void ", ""), r[12], float(r[14].replace(",", ""))
        per.setdefault(name, collections.defaultdict(list))[metric].append(val)
        if metric == "gpu__time_duration.sum":
            seq.append((name, r[8], val / 1e3))
    tot = sum(sum(m["gpu__time_duration.sum"]) for m in per.values())
    for name, m in per.items():
        t = m["gpu__time_duration.sum"]
        extra = "".join("  %s=%.3g" % (k.split("__")[-1], sum(v) / len(v)) for k, v in m.items() if k != "gpu__time_duration.sum")
        print("%-34s n=%4d avg=%9.1f us share=%5.1f%%%s" % (name[:34], len(t), sum(t) / len(t) / 1e3, 100 * sum(t) / tot, extra))
    for s in seq[-last:] if last else []:
        print("   %-34s grid %-16s %9.1f us" % s)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
