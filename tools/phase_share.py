"""Groups the per-line table of profiles/hot_lines.py by source ranges.  usage: python tools/phase_share.py table.txt name:lo-hi ..."""
import re
import sys
ranges = []
for a in sys.argv[2:]:
    n, r = a.split(":"); lo, hi = r.split("-"); ranges.append((n, int(lo), int(hi)))
acc = {n: [0, 0] for n, _, _ in ranges}; other = {}
for ln in open(sys.argv[1]):
    m = re.match(r'(\S+):(\d+)\s+(\d+)\s+[\d.]+%\s+(\d+)', ln)
    if not m:
        continue
    f, l, s, w = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))
    hit = False
    if f == 'observe.cuh':
        for n, a, b in ranges:
            if a <= l <= b:
                acc[n][0] += s; acc[n][1] += w; hit = True; break
    if not hit:
        o = other.setdefault(f, [0, 0]); o[0] += s; o[1] += w
acc.update(other)
ts = sum(v[0] for v in acc.values()); tw = sum(v[1] for v in acc.values())
for n, v in acc.items():
    print('%-32s samples %6d %5.1f%%  warp-inst %10d %5.1f%%' % (n, v[0], 100 * v[0] / ts, v[1], 100 * v[1] / tw))
