"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k1_|k2_" -c 300 --csv`)
into the text committed under profiles/: time and share per kernel, for the headline launches and for all of them.
usage: python tools/launch_list.py launches.csv [headline_launches=80] [frames_per_launch=32] > profiles/rN_launch_list.txt"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    n_head = int(sys.argv[2]) if len(sys.argv) > 2 else 80
    fpl = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        ns = float(r['Metric Value'].replace(',', ''))
        if r['Metric Unit'] == 'us':
            ns *= 1e3
        rows.append((r['Kernel Name'].split('(CUtensorMap_st')[0].split('(K')[0], ns / 1e3))

    def table(sel, per_frame):
        agg = OrderedDict()
        for k, us in sel:
            c = agg.setdefault(k, [0, 0.0])
            c[0] += 1
            c[1] += us
        tot = sum(c[1] for c in agg.values())
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            tail = '  (%.1f us per frame)' % (us / (n * fpl)) if per_frame else ''
            print('  %-120s %4d launches %10.1f us  %5.1f%%%s' % (k, n, us, 100 * us / tot, tail))

    print('headline (first %d launches): %.2f ms' % (n_head, sum(us for _, us in rows[:n_head]) / 1e3))
    table(rows[:n_head], True)
    print()
    print('all %d launches: %.2f ms' % (len(rows), sum(us for _, us in rows) / 1e3))
    table(rows, False)


if __name__ == '__main__':
    main()
