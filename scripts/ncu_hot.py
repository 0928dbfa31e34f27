"""Hot regions of one ncu --set full report: SASS instructions with their stall samples, grouped into contiguous
windows around every instruction above a threshold.  Usage: python scripts/ncu_hot.py rep.ncu-rep [min_share_pct] [context]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 3
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; data = rows[2:]
isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[isamp]) for r in data)
hot = [i for i, r in enumerate(data) if 100.0 * int(r[isamp]) / tot >= thr]
show = sorted(set(j for i in hot for j in range(max(0, i - ctx), min(len(data), i + ctx + 1))))
prev = None
for j in show:
    if prev is not None and j != prev + 1:
        print('   ...')
    r = data[j]
    st = sorted(((int(r[i]), n[6:]) for i, n in stall_cols if r[i] not in ('', '0')), reverse=True)[:2]
    print('%5d %6.2f%% ex=%-9s %-70s %s' % (j, 100.0 * int(r[isamp]) / tot, r[iex], r[isrc].strip()[:70], ' '.join('%s:%d' % (n, c) for c, n in st)))
    prev = j
