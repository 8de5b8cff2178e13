"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == 'Line No'][0]
hdr = rows[hi]
ci = hdr.index('Instructions Executed'); cs = hdr.index('# Samples')
data = []
for r in rows[hi + 1:]:
    if r[0] == '' or len(r) <= ci: continue
    try: data.append((int(r[ci]), int(r[cs]), int(r[0]), r[1][:120]))
    except Exception: pass
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print('total inst', tot, 'samples', tots)
for d in sorted(data, key=lambda x: -x[1])[:top]:
    print(f"{100*d[0]/tot:5.1f}% inst {100*d[1]/tots:5.1f}% smp  L{d[2]:>5} {d[3]}")
