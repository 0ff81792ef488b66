#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` per CUDA source line.

    python tools/ncu_lines.py src.csv [top_n]
Prints, for the hottest source lines: share of executed warp instructions, share of stall samples,
average active threads per instruction, and the dominant stall reasons.
"""
import csv
import sys
from collections import defaultdict


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and r[0] != "":  # a CUDA source line with aggregated metrics
            d = dict(zip(hdr[4:], r[4:]))
            lines.append((cur_file, r[0], r[1], d))
    tot_inst = sum(num(d["Instructions Executed"]) for *_, d in lines)
    tot_thr = sum(num(d["Thread Instructions Executed"]) for *_, d in lines)
    tot_smp = sum(num(d["# Samples"]) for *_, d in lines)
    print("total warp inst %.4e  thread inst %.4e (%.2f thr/inst)  samples %d" % (tot_inst, tot_thr, tot_thr / tot_inst, tot_smp))
    stall_keys = [k for k in lines[0][3] if k.startswith("stall_") and "Not Issued" not in k]
    agg = defaultdict(float)
    for *_, d in lines:
        for k in stall_keys:
            agg[k] += num(d[k])
    print("stall mix: " + "  ".join("%s %.1f%%" % (k[6:], 100 * v / tot_smp) for k, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
    byfile = defaultdict(lambda: [0.0, 0.0])
    for f, _, _, d in lines:
        byfile[f][0] += num(d["Instructions Executed"])
        byfile[f][1] += num(d["# Samples"])
    for f, (i, s) in byfile.items():
        print("  %-18s inst %5.1f%%  samples %5.1f%%" % (f, 100 * i / tot_inst, 100 * s / tot_smp))
    for f, ln, src, d in sorted(lines, key=lambda t: -num(t[3]["# Samples"]))[:top]:
        st = sorted(((num(d[k]), k[6:]) for k in stall_keys), reverse=True)[:3]
        print("%5.2f%%i %5.2f%%s thr %4.1f | %-14s:%-4s | %-70s | %s" % (
            100 * num(d["Instructions Executed"]) / tot_inst, 100 * num(d["# Samples"]) / tot_smp,
            num(d["Thread Instructions Executed"]) / max(1.0, num(d["Instructions Executed"])), f, ln, src.strip()[:70],
            " ".join("%s:%d" % (k, v) for v, k in st if v > 0)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
