#!/usr/bin/env python
"""Per-trip kernel times of the LAST complete solve in an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
    python tools/launch_table.py gpurun_out/xxx_launches.csv [which_solve]"""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    seq = []
    for row in csv.DictReader(lines):
        if row['Metric Name'] != 'gpu__time_duration.sum':
            continue
        name = row['Kernel Name'].split('<')[0].split('(')[0].replace('void ', '').replace('ilqr::', '')
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        seq.append((name, v))
    return seq


def main(path, which=None):
    seq = load(path)
    starts = [i for i, s in enumerate(seq) if s[0] == 'phase_begin_kernel']
    i0 = starts[which if which is not None else (-2 if len(starts) > 1 else -1)]
    end = min([i for i in starts if i > i0] + [len(seq)])
    trips, cur = [], None
    for name, v in seq[i0 + 1:end]:
        if not name.startswith('phase_'):
            continue
        if name in ('phase_sweep_kernel', 'phase_pre_warp_kernel') or cur is None:
            cur = collections.OrderedDict()
            trips.append(cur)
        cur[name.replace('phase_', '').replace('_kernel', '')] = v
    tot = collections.defaultdict(float)
    for n, t in enumerate(trips):
        for k, v in t.items():
            tot[k] += v
        if n < 4 or n % 8 == 0 or n >= len(trips) - 3:
            print("%3d  " % n + "  ".join("%s %7.1f" % kv for kv in t.items()) + "   sum %7.1f" % sum(t.values()))
    print(len(trips), "trips;", "  ".join("%s %.1f ms" % (k, v / 1e3) for k, v in tot.items()), "; total %.1f ms" % (sum(tot.values()) / 1e3))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
