"""Hot source lines of a kernel from an .ncu-rep captured with --import-source on (-lineinfo build).

usage: python scripts/ncu_hot_lines.py REPORT.ncu-rep [top_n]
Runs `ncu -i REPORT --page source --csv --print-source cuda,sass` and aggregates the warp-stall samples and
executed instructions per (file, line); inlined headers show up under their own file names.
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO("\n".join(l for l in out.splitlines() if not l.startswith("==")))))
cur_file, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples")
        i_i = hdr.index("Instructions Executed")
        i_long = hdr.index("stall_long_sb")
        i_wait = hdr.index("stall_wait")
        i_short = hdr.index("stall_short_sb")
        continue
    if hdr is None or r[0] == "":   # SASS rows carry an empty line number
        continue
    try:
        key = (cur_file, int(r[0]))
        agg[key] = (r[1].strip()[:100], int(r[i_s] or 0), int(r[i_i] or 0), int(r[i_long] or 0), int(r[i_wait] or 0), int(r[i_short] or 0))
    except (ValueError, IndexError):
        pass
tot = sum(v[1] for v in agg.values()) or 1
print(f"total samples {tot}")
print(f"{'file:line':28s} {'samp%':>6s} {'inst':>9s} {'long':>5s} {'wait':>5s} {'short':>5s}  source")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f + ':' + str(ln):28s} {100 * v[1] / tot:6.1f} {v[2]:9d} {v[3]:5d} {v[4]:5d} {v[5]:5d}  {v[0]}")
