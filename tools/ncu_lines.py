#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump: instructions executed and
stall samples by file:line (top N).  Usage: tools/ncu_lines.py dump.csv [N]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fname = None
    agg = []
    tot_inst = tot_samp = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if len(r) > 7 and r[2] == "-" and r[0].isdigit():      # a source line row (aggregated over its SASS)
            try:
                samp, inst = int(r[4] or 0), int(r[7] or 0)
            except ValueError:
                continue
            agg.append((inst, samp, fname, int(r[0]), r[1].strip()[:110]))
            tot_inst += inst
            tot_samp += samp
    print("total inst %d, samples %d" % (tot_inst, tot_samp))
    for inst, samp, f, ln, src in sorted(agg, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% stall  %-16s:%-4d %s" % (100.0 * inst / max(tot_inst, 1), 100.0 * samp / max(tot_samp, 1), f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
