#!/usr/bin/env python
"""Aggregate an ncu per-launch metric list (--csv) over the kernels of ONE solve (the last complete one in the file).

    python tools/ncu_solve_summary.py launches.csv [out.json]
Expects the metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum, smsp__inst_executed.sum,
smsp__thread_inst_executed.sum, sm__inst_executed_pipe_fp64.sum (any subset).  A solve = the launches from the
phase_begin kernel (or the first ilqr_warp_kernel after an init launch) to the launch before the next init."""
import collections
import csv
import json
import sys

UNITS = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1.0, "": 1.0}


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        lid = int(row["ID"])
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("ilqr::", "")
        d = launches.setdefault(lid, {"name": name})
        v = float(row["Metric Value"].replace(",", "")) * UNITS.get(row["Metric Unit"], 1.0)
        d[row["Metric Name"]] = v
    return list(launches.values())


def main(path, out=None):
    L = load(path)
    # solves are delimited by the init launch (ilqr_warp_kernel right after which phase_begin follows) -> use phase_begin
    begins = [i for i, l in enumerate(L) if l["name"].startswith("phase_begin")]
    if begins:
        i0, i1 = begins[-1], len(L)  # the last solve: from its phase_begin to the end of the capture
    else:
        i0, i1 = 0, len(L)
    solve = L[i0:i1]
    per = collections.OrderedDict()
    for l in solve:
        k = l["name"].split("<")[0]
        a = per.setdefault(k, collections.defaultdict(float))
        a["launches"] += 1
        for m, v in l.items():
            if m != "name":
                a[m] += v
    tot = collections.defaultdict(float)
    for a in per.values():
        for m, v in a.items():
            tot[m] += v
    res = {"kernels": {k: dict(v) for k, v in per.items()}, "total": dict(tot)}
    t = tot.get("gpu__time_duration.sum", 0.0)
    print("one solve: %d launches, %.2f ms of kernel time (serialised, cold caches)" % (len(solve), t * 1e3))
    for k, a in per.items():
        line = "  %-28s n=%4d  %8.2f ms (%4.1f %%)" % (k, a["launches"], a.get("gpu__time_duration.sum", 0) * 1e3,
                                                   100 * a.get("gpu__time_duration.sum", 0) / max(t, 1e-30))
        if "dram__bytes_read.sum" in a:
            line += "  dram r %.2f GB w %.2f GB" % (a["dram__bytes_read.sum"] / 1e9, a["dram__bytes_write.sum"] / 1e9)
        if "smsp__inst_executed.sum" in a:
            line += "  warp-inst %.3e thr/inst %.1f" % (a["smsp__inst_executed.sum"], a.get("smsp__thread_inst_executed.sum", 0) /
                                                         max(a["smsp__inst_executed.sum"], 1))
        print(line)
    if "dram__bytes_read.sum" in tot:
        print("  dram total: read %.2f GB + write %.2f GB = %.2f GB" % (tot["dram__bytes_read.sum"] / 1e9, tot["dram__bytes_write.sum"] / 1e9,
                                                                       (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / 1e9))
    if "smsp__inst_executed.sum" in tot:
        print("  warp instructions %.4e, threads per instruction %.2f" % (tot["smsp__inst_executed.sum"],
                                                                          tot.get("smsp__thread_inst_executed.sum", 0) / tot["smsp__inst_executed.sum"]))
    if out:
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
