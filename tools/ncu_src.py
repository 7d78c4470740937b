"""Aggregate an `ncu --page source --csv` dump: dynamic instruction mix, stall samples by opcode and hottest SASS lines."""
import csv, sys, re, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stalls = collections.Counter(); lines = []
tot = 0; tots = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix['Source']]
    try: n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    except ValueError: continue
    m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', src)
    op = m.group(2) if m else '?'
    ops[op] += n; stalls[op] += s; tot += n; tots += s
    lines.append((s, n, r[ix['Address']], src))
print('total warp instructions', tot, 'samples', tots)
print('%-10s %12s %6s %8s %6s' % ('op', 'executed', '%', 'samples', '%'))
for op, n in ops.most_common(22):
    print('%-10s %12d %6.1f %8d %6.1f' % (op, n, 100.0 * n / tot, stalls[op], 100.0 * stalls[op] / max(tots, 1)))
print('--- hottest lines by stall samples')
for s, n, a, src in sorted(lines, reverse=True)[:top]:
    print('%7d %9d %s %s' % (s, n, a[-5:], src[:110]))
