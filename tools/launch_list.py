#!/usr/bin/env python
"""Per-kernel summary of one step out of an `ncu --metrics gpu__time_duration.sum --csv` launch list
(usage: launch_list.py list.csv [first kernel of a step = batch_unpack_kernel] [which step])."""
import collections
import csv
import sys


def load(p):
    rows = list(csv.reader(open(p)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hi]
    return [dict(zip(h, r)) for r in rows[hi + 2:] if len(r) >= len(h)]


def main():
    rows = load(sys.argv[1])
    first = sys.argv[2] if len(sys.argv) > 2 else 'batch_unpack_kernel'
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    names = [r['Kernel Name'].split('(')[0].replace('unnamed>::', '').replace('void ', '') for r in rows]
    starts = [i for i, n in enumerate(names) if n == first] + [len(rows)]
    s0, s1 = starts[which], starts[which + 1]
    tot = 0; agg = collections.OrderedDict()
    for r, n in zip(rows[s0:s1], names[s0:s1]):
        t = float(r['Metric Value'].replace(',', '')) / 1000; tot += t
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
        print(f"{t:9.1f} us  {n[:50]:50s} grid {r['Grid Size']} blk {r['Block Size']}")
    print('total us', round(tot, 1), 'launches', s1 - s0)
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.1f} us {100 * t / tot:5.1f}%  x{c:3d} {n}")


if __name__ == "__main__":
    main()
