"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv) into SASS segments:
share of warp instructions, average active threads, share of stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
kern, cur, hdr = {}, None, None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = r[1]; kern[cur] = []; continue
    if r and r[0] == 'Address':
        hdr = r; continue
    if cur and r:
        kern[cur].append(r)
for k, rs in kern.items():
    iE = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iS = hdr.index('Source'); iSm = hdr.index('# Samples')
    tot = sum(int(r[iE]) for r in rs); totT = sum(int(r[iT]) for r in rs); totS = max(1, sum(int(r[iSm]) for r in rs))
    print('=====', k[:50], 'sass', len(rs), 'warp inst %.3e' % tot, 'avg thr %.2f' % (totT / max(tot, 1)))
    seg = []
    for i, r in enumerate(rs):
        e = int(r[iE]); t = int(r[iT]); sm = int(r[iSm])
        if seg and abs(seg[-1]['e0'] - e) <= 0.15 * max(seg[-1]['e0'], e, 1):
            g = seg[-1]; g['n'] += 1; g['e'] += e; g['t'] += t; g['s'] += sm; g['end'] = i
        else:
            seg.append({'start': i, 'end': i, 'n': 1, 'e0': e, 'e': e, 't': t, 's': sm})
    for g in seg:
        if g['e'] < min_share / 100 * tot:
            continue
        ops = {}
        for r in rs[g['start']:g['end'] + 1]:
            f = r[iS].split()
            op = f[1] if f[0].startswith('@') else f[0]
            op = op.split('.')[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda x: -x[1])[:7]
        print('seg %5d-%5d n=%4d exec=%11d share=%5.1f%% thr=%5.1f stall=%5.1f%% %s' % (
            g['start'], g['end'], g['n'], g['e0'], 100 * g['e'] / tot, g['t'] / max(g['e'], 1), 100 * g['s'] / totS, top))
