"""Split a kernel's warp-stall samples and executed instructions into the code segments between CTA barriers
(development aid; reads a .ncu-rep captured with --import-source on)."""
import csv, subprocess, collections, sys
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r][0]
h = rows[hi]; si = h.index('Source'); ei = h.index('Instructions Executed'); wi = h.index('Warp Stall Sampling (All Samples)')
data = []
for r in rows[hi + 1:]:
    if len(r) <= ei: continue
    try: n = int(r[ei]); s = int(r[wi])
    except ValueError: continue
    data.append((s, n, r[si]))
tot = sum(d[0] for d in data); totn = sum(d[1] for d in data)
seg = 0; segs = collections.OrderedDict(); first = {}; ops = collections.defaultdict(collections.Counter)
for i, (s, n, t) in enumerate(data):
    segs.setdefault(seg, [0, 0, 0]); segs[seg][0] += s; segs[seg][1] += n; segs[seg][2] += 1
    first.setdefault(seg, i)
    tt = t.split(); op = tt[0] if not tt[0].startswith('@') else tt[1]
    ops[seg][op.split('.')[0]] += n
    if 'BAR.SYNC' in t: seg += 1
print("total samples", tot, "warp instructions", totn)
for k, v in segs.items():
    if v[0] > 0.01 * tot:
        top = ", ".join("%s %.0f%%" % (o, 100 * c / max(1, v[1])) for o, c in ops[k].most_common(5))
        print("segment %3d: samples %5.1f%%  instrs %5.1f%% (%9d, static %4d) | %s" % (k, 100 * v[0] / tot, 100 * v[1] / totn, v[1], v[2], top))
