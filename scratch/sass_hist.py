"""opcode histogram + hottest instructions of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv)"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ie, src, smp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
c = collections.Counter(); s = collections.Counter(); tot = 0; stot = 0
lines = []
for r in rows[2:]:
    try: n = int(r[ie]); k = int(r[smp])
    except Exception: continue
    t = r[src].split()
    op = t[1] if t[0].startswith('@') else t[0]
    c[op] += n; s[op] += k; tot += n; stot += k
    lines.append((n, k, r[src].strip()))
print("total warp-instructions", tot, "samples", stot)
for k, v in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{k:28s} {v:10d} {100*v/tot:5.1f}%   samples {100*s[k]/max(stot,1):5.1f}%")
print("--- hottest by samples")
for n, k, t in sorted(lines, key=lambda x: -x[1])[:15]:
    print(f"{k:6d} {n:9d}  {t}")
