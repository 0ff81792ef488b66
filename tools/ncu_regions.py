#!/usr/bin/env python
"""Attribute ncu source-page samples/instructions to solver REGIONS by SASS address.

Inlined callees (trig.cuh, models.cuh, boxqp.cuh) have no caller in the per-line view; but every SASS
instruction has an address, and the code of each region (derivative sweep, backward step, candidate
rollout ...) is contiguous.  Walk the instructions in address order and label each with the region of
the nearest ilqr_core.cuh line around it.

    python tools/ncu_regions.py src_cuda_sass.csv path/to/ilqr_core.cuh
"""
import bisect
import csv
import re
import sys
from collections import defaultdict


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def regions_of(core_path):
    """line ranges of the member functions of Core"""
    out = []
    pat = re.compile(r"^  ILQR_HD (?:static )?[\w:<>&\* ]+?(\w+)\(")
    for i, line in enumerate(open(core_path), 1):
        m = pat.match(line)
        if m:
            out.append((i, m.group(1)))
    return out


def main(path, core):
    regs = regions_of(core)
    starts = [r[0] for r in regs]
    rows = list(csv.reader(open(path)))
    cur_file, cur_line, hdr = None, None, None
    inst = []  # (addr, file, line, n_inst, samples)
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name",):
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] != "":
            cur_line = int(r[0])
            continue
        if not r[2].startswith("0x"):
            continue
        inst.append((int(r[2], 16), cur_file, cur_line, num(r[7]), num(r[6])))
    inst.sort()
    # label: nearest ilqr_core.cuh instruction in address order (looking both ways)
    core_idx = [i for i, t in enumerate(inst) if t[1] == "ilqr_core.cuh"]
    agg = defaultdict(lambda: [0.0, 0.0])
    for i, (addr, f, ln, n, s) in enumerate(inst):
        j = bisect.bisect_left(core_idx, i)
        cands = [core_idx[k] for k in (j - 1, j) if 0 <= k < len(core_idx)]
        k = min(cands, key=lambda c: abs(c - i))
        cl = inst[k][2]
        name = regs[max(0, bisect.bisect_right(starts, cl) - 1)][1] if regs else "?"
        agg[name][0] += n
        agg[name][1] += s
    ti = sum(v[0] for v in agg.values())
    ts = sum(v[1] for v in agg.values())
    print("region                      inst%   samples%")
    for name, (n, s) in sorted(agg.items(), key=lambda t: -t[1][1]):
        print("%-26s %6.1f %9.1f" % (name, 100 * n / ti, 100 * s / ts))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
