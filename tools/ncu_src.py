#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump: share of executed warp
instructions, share of stall samples and active lanes per line (usage: ncu_src.py dump.csv [top_n] [file filter])."""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    hdr = None; cur = None; agg = collections.OrderedDict(); files = collections.Counter()
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if len(r) >= 2 and r[0] == "Function Name": continue
        if r and r[0] == "Line No": hdr = r; continue
        if hdr is None or len(r) < len(hdr): continue
        try:
            ln = int(r[0]); inst = int(r[hdr.index("Instructions Executed")] or 0)
            samp = int(r[hdr.index("# Samples")] or 0); thr = int(r[hdr.index("Thread Instructions Executed")] or 0)
        except ValueError:
            continue
        a = agg.setdefault((cur, ln, r[1].strip()[:90]), [0, 0, 0]); a[0] += inst; a[1] += samp; a[2] += thr
        files[cur] += inst
    T = sum(v[0] for v in agg.values()) or 1; S = sum(v[1] for v in agg.values()) or 1
    print("warp instructions", T, "samples", S, dict(files))
    n = 0
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if filt and filt not in k[0]: continue
        print("%5.1f%% inst %5.1f%% samp  lanes %4.1f | %s:%d %s" % (100 * v[0] / T, 100 * v[1] / S, v[2] / max(v[0], 1), k[0], k[1], k[2]))
        n += 1
        if n >= top: break


if __name__ == "__main__":
    main()
