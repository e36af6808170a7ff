"""Dynamic SASS opcode mix from an `ncu --page source --csv` export."""
import csv, sys, collections
f = sys.argv[1]
rows = list(csv.reader(open(f)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ie, src, smp = ci['Instructions Executed'], ci['Source'], ci['# Samples']
mix, stall = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= ie: continue
    s = r[src].strip()
    if s.startswith('@'): s = s.split(None, 1)[1]
    op = s.split()[0].split('.')[0] if s else '?'
    n = int(r[ie] or 0); mix[op] += n; tot += n; stall[op] += int(r[smp] or 0)
warps = int(rows[2][ie])
print("total warp-instr %d, per warp %.0f" % (tot, tot / warps))
for op, n in mix.most_common(28):
    print("  %-10s %8.1f per warp  %5.1f%%   samples %5.1f%%" % (op, n / warps, 100 * n / tot, 100 * stall[op] / max(1, sum(stall.values()))))
