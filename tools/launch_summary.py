"""Per-kernel totals of an ncu launch list (`ncu --profile-from-start off --metrics gpu__time_duration.sum
--clock-control none --csv --log-file X.csv python bench.py --steps K ...`): launches, total time, share of the listed
time.  The per-launch times are cold-cache and serialised, so only the SHARES compare with bench.py's live numbers.
usage: python tools/launch_summary.py X.csv [steps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors='replace')))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[start]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui].strip(), 1e-3)
    a = agg.setdefault(r[ki].split('(')[0][:72], [0, 0.0])
    a[0] += 1
    a[1] += v * scale
tot = sum(a[1] for a in agg.values())
print('%d launches, %.1f us listed (%d timed step(s): %.1f launches and %.3f ms per step)' %
      (sum(a[0] for a in agg.values()), tot, steps, sum(a[0] for a in agg.values()) / steps, tot / steps / 1e3))
print('%-72s %8s %12s %10s %7s' % ('kernel', 'launches', 'total us', 'us/launch', 'share'))
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print('%-72s %8d %12.1f %10.2f %6.1f%%' % (n, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
mine = sum(a[1] for n, a in agg.items() if 'gwtf::' in n)
print('kernels of libgwtf.so: %.1f%% of the listed time, %d launches' %
      (100 * mine / tot, sum(a[0] for n, a in agg.items() if 'gwtf::' in n)))
