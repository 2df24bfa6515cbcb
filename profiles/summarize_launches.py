"""Median per-kernel metrics of an `ncu --csv --log-file` launch list (any --metrics set)."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
data = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows:
    if len(r) > 5 and r[0] == 'ID':
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        data[d['Kernel Name'].split('(')[0][-40:]][d['Metric Name'] + ' ' + d['Metric Unit']].append(v)
for k, m in data.items():
    print(k)
    for n, v in m.items():
        print('   %-62s n=%-3d median=%.4g' % (n, len(v), sorted(v)[len(v) // 2]))
