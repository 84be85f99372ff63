"""Top stalled SASS instructions of an `ncu --page source --csv` dump (developer tool)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
tot = sum(int(r[ix['# Samples']] or 0) for r in body)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter()
for r in body:
    for c in stall_cols:
        agg[c] += int(r[ix[c]] or 0)
print("total samples", tot, "instructions", len(body))
print("stall totals:", ", ".join("%s %.1f%%" % (k.replace('stall_', ''), 100 * v / tot) for k, v in agg.most_common(8)))
top = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]
for i in top:
    r = body[i]
    st = sorted(((int(r[ix[c]] or 0), c.replace('stall_', '')) for c in stall_cols), reverse=True)[:2]
    print("%5d %-58s %6s (%.1f%%) %s" % (i, r[ix['Source']].strip()[:58], r[ix['# Samples']], 100 * int(r[ix['# Samples']]) / tot, st))
