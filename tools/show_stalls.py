#!/usr/bin/env python
"""Print the per-launch warp-stall table of an `ncu --section WarpStateStats --section SchedulerStats
--csv --page raw` log of tools/move_breakdown.py (launches: all, then warm-up + timed per move type)."""
import csv
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    names = ['all', 'crank_w', 'crank', 'pivot_w', 'pivot', 'slide_w', 'slide', 'tang_w', 'tang', 'bind_w', 'bind']
    keys = [k for k in hdr if 'issue_stalled' in k and 'per_issue_active' in k and 'not_issued' not in k]
    print(path)
    print('%-9s' % '', ' '.join('%6s' % k.split('stalled_')[1].split('_per')[0][:6] for k in keys), '  lat issue%')
    for i, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        if i >= len(names) or names[i].endswith('_w'):
            continue
        print('%-9s' % names[i], ' '.join('%6.2f' % float(d[k]) for k in keys),
              '%5.1f' % float(d['smsp__average_warp_latency_per_inst_issued.ratio']),
              '%5.1f' % float(d['smsp__issue_active.avg.pct_of_peak_sustained_active']))
