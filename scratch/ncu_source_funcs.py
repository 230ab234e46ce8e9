"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per enclosing C++
function (nearest preceding line that looks like a function or struct-method header)."""
import csv, collections, re, sys
path = sys.argv[1]
only = sys.argv[2] if len(sys.argv) > 2 else None
hdr = None; cur = None
rows = []
for r in csv.reader(open(path)):
    if len(r) == 2 and r[0] in ('File Path', 'File Name'):
        cur = r[1]; continue
    if len(r) > 6 and r[0] == 'Line No':
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != '':
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            rows.append((cur, int(r[0]), int(d['# Samples']), int(d['Instructions Executed']),
                         int(d['Thread Instructions Executed'])))
        except ValueError:
            pass
func_re = re.compile(r'^\s*(?:template<[^>]*>\s*)?(?:B2_D|B2_HD|B2_NOINLINE|B2_UNIV_FN|B2_GEO_FN|__global__)\b.*?(\w+)\s*\(')
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = sum(x[2] for x in rows); toti = sum(x[3] for x in rows)
cache = {}
for f, line, samp, inst, thr in rows:
    if only and only not in f:
        continue
    if f not in cache:
        names = []
        try:
            for i, text in enumerate(open(f), 1):
                m = func_re.match(text)
                if m:
                    names.append((i, m.group(1)))
        except OSError:
            pass
        cache[f] = names
    name = '?'
    for i, n in cache[f]:
        if i <= line:
            name = n
        else:
            break
    a = agg[(f.split('/')[-1], name)]
    a[0] += samp; a[1] += inst; a[2] += thr
for (f, n), (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print('%5.2f%% samples %5.2f%% inst  lanes %4.1f  %s:%s' % (100 * s / tot, 100 * i / toti, t / max(i, 1), f, n))
