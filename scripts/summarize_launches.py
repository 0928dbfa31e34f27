"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
Usage: python scripts/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/<name>.md"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        rows.append((int(r['ID']), r['Kernel Name'].split('(')[0], r['Grid Size'], r['Block Size'],
                     float(r['Metric Value'].replace(',', '')), r['Metric Unit']))
    rows = [r for r in rows if r[0] >= skip]
    agg = OrderedDict()
    for _, name, grid, block, v, unit in rows:
        ns = v * {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1; a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print('| kernel | launches | total ms | share | avg us | grid | block |')
    print('|---|---|---|---|---|---|---|')
    for name, (n, ns, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.3f | %.1f%% | %.1f | %s | %s |' % (name, n, ns / 1e6, 100 * ns / tot, ns / n / 1e3, grid, block))
    print('\ntotal %.3f ms over %d launches (ncu-serialised, cold-cache: compare shares, not absolutes)' % (tot / 1e6, len(rows)))


if __name__ == '__main__':
    main()
