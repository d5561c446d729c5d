#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source sass,cuda` output by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python src_hotspots.py src.csv [N]"""
import collections
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    agg = collections.defaultdict(lambda: [0, 0])
    cur, hdr = None, None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur, hdr = r[1].split('/')[-1], None
            continue
        if len(r) == 2:
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        key = (cur, int(r[0]), r[1][:100].strip())
        agg[key][0] += num(d.get("# Samples"))
        agg[key][1] += num(d.get("Instructions Executed"))
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print(f"total stall samples {ts}, warp instructions {ti}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0] * 100 / ts:5.1f}% samples {v[1] * 100 / ti:5.1f}% inst  {k[0]}:{k[1]}  {k[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
