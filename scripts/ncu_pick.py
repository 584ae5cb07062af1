"""Print selected raw metrics (substring filters from argv) of every kernel in an `ncu --page raw --csv` dump on stdin."""
import csv, sys
pats = sys.argv[1:] or ["op_red", "stalled"]
rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("----", d["Kernel Name"][:60])
    for k in hdr:
        if any(p in k for p in pats):
            print(f"   {k:90s} {d[k]}")
