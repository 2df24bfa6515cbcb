"""Per-source-line instruction counts of one kernel from an ncu report with -lineinfo / --import-source on:
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv | python profiles/ncu_lines.py [min_total_inst] [top_n]"""
import collections, csv, sys
rows = list(csv.reader(sys.stdin))
min_total = float(sys.argv[1]) if len(sys.argv) > 1 else 1e6
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
secs, cur = [], None
for r in rows:
    if len(r) > 3 and r[0] == 'Line No':
        cur = {'hdr': r, 'rows': []}; secs.append(cur); continue
    if cur is not None and len(r) == len(cur['hdr']):
        cur['rows'].append(r)
for s in secs:
    h = s['hdr']; iI = h.index('Instructions Executed'); iT = h.index('Thread Instructions Executed'); iS = h.index('# Samples')
    agg = collections.defaultdict(lambda: [0, 0, 0, ''])
    for r in s['rows']:
        try:
            ln, ins, thr, sm = int(r[0]), int(r[iI]), int(r[iT]), int(r[iS])
        except ValueError:
            continue
        a = agg[ln]; a[0] += ins; a[1] += thr; a[2] += sm; a[3] = r[1]
    tot = sum(a[0] for a in agg.values())
    if tot < min_total:
        continue
    print('section: %d warp instructions, %d samples' % (tot, sum(a[2] for a in agg.values())))
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%5d inst %9d (%4.1f%%) lanes %4.1f samples %5d | %s' % (ln, a[0], 100 * a[0] / tot, a[1] / max(a[0], 1), a[2], a[3].strip()[:100]))
