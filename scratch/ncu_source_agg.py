"""Aggregate `ncu --page source --print-source cuda,sass --csv` per file, function-ish region and line."""
import csv, collections, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None
per_file = collections.defaultdict(lambda: [0, 0, 0.0])
per_line = []
stall_cols = None
for r in csv.reader(open(path)):
    if len(r) == 2 and r[0] in ('File Path', 'File Name'):
        cur = r[1]; continue
    if len(r) > 6 and r[0] == 'Line No':
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != '':
        # a CUDA source line row (SASS rows have empty Line No)
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            samp = int(d['# Samples']); inst = int(d['Instructions Executed'])
            thr = int(d['Thread Instructions Executed'])
        except ValueError:
            continue
        per_file[cur][0] += samp; per_file[cur][1] += inst; per_file[cur][2] += thr
        stalls = {k: int(v) for k, v in d.items() if k.startswith('stall_') and '(' not in k and v.isdigit() and int(v)}
        per_line.append((samp, inst, cur.split('/')[-1], r[0], r[1].strip()[:100], stalls))
tot = sum(v[0] for v in per_file.values()); toti = sum(v[1] for v in per_file.values())
print('total samples', tot, 'warp instructions', toti)
for k, v in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
    print('%6.2f%% samples %6.2f%% inst  lanes %.1f  %s' % (100 * v[0] / tot, 100 * v[1] / toti, v[2] / max(v[1], 1), k))
per_line.sort(key=lambda x: -x[0])
for l in per_line[:top]:
    st = sorted(l[5].items(), key=lambda kv: -kv[1])[:3]
    print('%5.2f%% %5.2f%% %s:%s  %s   %s' % (100 * l[0] / tot, 100 * l[1] / toti, l[2], l[3], l[4], st))
