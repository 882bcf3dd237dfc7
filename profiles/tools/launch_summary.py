"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals for one training
iteration.  usage: python profiles/tools/launch_summary.py <launches.csv> <first_row> <last_row> [label]
Rows are 0-based indices into the launch list (pick one steady-state iteration; see profiles/README.md)."""
import collections
import csv
import re
import sys


def load(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    for x in csv.DictReader(lines):
        try:
            rows.append((x['Kernel Name'], float(x['Metric Value'].replace(',', '')), x['Grid Size']))
        except (KeyError, ValueError):
            pass
    return rows


def short(n):
    n = re.sub(r'<unnamed>::|\(anonymous namespace\)::', '', n)
    return re.sub(r'\(.*', '', n).replace('void ', '')[:72]


def main():
    rows = load(sys.argv[1])
    a, b = int(sys.argv[2]), int(sys.argv[3])
    label = sys.argv[4] if len(sys.argv) > 4 else ''
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t, g in rows[a:b]:
        agg[short(n)][0] += 1
        agg[short(n)][1] += t
    tot = sum(v[1] for v in agg.values())
    print('# %s rows [%d, %d) of %s: %d launches, %.1f us of kernel time (cold-cache, serialised by ncu)' % (label, a, b, sys.argv[1], b - a, tot / 1e3))
    w = csv.writer(sys.stdout)
    w.writerow(['kernel', 'launches', 'total_us', 'share_pct'])
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, v[0], '%.1f' % (v[1] / 1e3), '%.2f' % (100 * v[1] / tot)])


if __name__ == '__main__':
    main()
