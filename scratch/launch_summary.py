"""Share of device time per kernel from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv, collections, sys, gzip
path = sys.argv[1]
op = gzip.open if path.endswith('.gz') else open
rows = [r for r in csv.reader(op(path, 'rt')) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = r[ik].split('(')[0].replace('void ', '')
    agg[name][0] += 1
    agg[name][1] += float(r[iv].replace(',', ''))
tot = sum(v[1] for v in agg.values())
print('%d launches, %.1f ms of kernel time (cold-cache, serialised)' % (sum(v[0] for v in agg.values()), tot / 1e6))
print('| kernel | launches | mean us | share |\n|---|---|---|---|')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('| %s | %d | %.1f | %.1f %% |' % (k, n, t / n / 1e3, 100 * t / tot))
