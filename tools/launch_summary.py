"""Summarise an ncu launch list (gpu__time_duration per launch) of one forward pass: python tools/launch_summary.py gpurun_out/launches.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value'); gs = hdr.index('Grid Size')
data = [(r[kn], float(r[mv].replace(',', '')), r[gs]) for r in rows[hi + 1:] if len(r) > mv and r[mv].replace(',', '').replace('.', '').isdigit()]
short = lambda n: re.sub(r'\(.*', '', n).replace('void ', '').replace('rb::', '')[:60]
# one forward = from the stem input kernel of fnet to the next one that follows an upsample kernel
starts = [i for i, d in enumerate(data) if 'enc_stem_s2d' in d[0] or 'enc_gather_img8' in d[0]]  # stem of fnet / cnet (r01: the gather)
ups = [i for i, d in enumerate(data) if 'upsample_convex' in d[0] or 'upflow8' in d[0]]
fw = None
for u in ups:
    s = [x for x in starts if x < u]
    if len(s) >= 2:
        fw = (s[-2], u + 1)
if fw is None:
    fw = (0, len(data))
seg = data[fw[0]:fw[1]]
print(f'forward pass: launches {fw[0]}..{fw[1]} ({len(seg)} kernels), {sum(d[1] for d in seg)/1e3:.1f} us (cold-cache, serialised)')
first_lookup = next((i for i, d in enumerate(seg) if 'corr_lookup' in d[0]), len(seg))
phases = [('encoders+corr build', seg[:first_lookup]), ('iterations+tail', seg[first_lookup:])]
for name, part in phases:
    agg = collections.OrderedDict()
    for n, t, g in part:
        a = agg.setdefault(short(n), [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f'== {name}: {len(part)} launches, {tot/1e3:.1f} us')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f'   {v[1]/1e3:9.1f} us {v[0]:5d}x {100*v[1]/tot:5.1f}%  {k}')
if '-v' in sys.argv:
    for n, t, g in seg[:first_lookup]:
        print(f'   {t/1e3:8.1f} us {g:>16} {short(n)}')
    li = [i for i, d in enumerate(seg) if 'corr_lookup' in d[0]]
    if len(li) > 6:
        print('iteration 5:')
        for n, t, g in seg[li[5]:li[6]]:
            print(f'   {t/1e3:8.1f} us {g:>16} {short(n)}')
