"""Splits the SASS of one profiled kernel (ncu --set full --import-source on) into segments at the barriers and
prints, per segment, executed warp instructions, stall samples and the opcode mix (development aid)."""
import csv, collections, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r][0]
h = rows[hi]; si = h.index('Source'); ei = h.index('Instructions Executed'); wi = h.index('Warp Stall Sampling (All Samples)')
ai = h.index('Address') if 'Address' in h else None
segs = []; cur = dict(n=0, s=0, ops=collections.Counter(), first=None, lines=0, sops=collections.Counter())
tot = ts = 0
for r in rows[hi + 1:]:
    if len(r) <= ei: continue
    try: n = int(r[ei]); s = int(r[wi])
    except: continue
    t = r[si].split()
    op = (t[0] if not t[0].startswith('@') else t[1]).split('.')[0]
    if cur['first'] is None: cur['first'] = r[ai] if ai is not None else ''
    cur['n'] += n; cur['s'] += s; cur['ops'][op] += n; cur['sops'][op] += s; cur['lines'] += 1
    tot += n; ts += s
    if op in ('BAR', 'WARPSYNC'):
        segs.append(cur); cur = dict(n=0, s=0, ops=collections.Counter(), first=None, lines=0, sops=collections.Counter())
segs.append(cur)
print("total instr %d samples %d" % (tot, ts))
for i, g in enumerate(segs):
    if g['n'] == 0: continue
    mix = ' '.join('%s %.0f%%' % (k, 100.0 * v / g['n']) for k, v in g['ops'].most_common(6))
    smix = ' '.join('%s %.0f%%' % (k, 100.0 * v / max(1, g['s'])) for k, v in g['sops'].most_common(4))
    print("seg %2d @%s lines %5d instr %5.1f%% samples %5.1f%% | %s || stall: %s" % (i, g['first'][-6:], g['lines'], 100.0 * g['n'] / tot, 100.0 * g['s'] / ts, mix, smix))
