"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/summarize_launches.py gpurun_out/launches_b1.csv > profiles/<name>.txt"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1., 'ms': 1e3, 's': 1e6}.get(row['Metric Unit'], 1.)
        a = agg.setdefault(row['Kernel Name'][:70], [0, 0.])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print('# %s: %d launches, %.1f us total (ncu-serialised, cold cache: compare SHARES)' % (path, sum(a[0] for a in agg.values()), tot))
    print('%-70s %6s %12s %10s %7s' % ('kernel', 'n', 'total_us', 'avg_us', 'share'))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('%-70s %6d %12.1f %10.1f %7.3f' % (k, a[0], a[1], a[1] / a[0], a[1] / tot))


if __name__ == '__main__':
    main(sys.argv[1])
