"""Dynamic opcode histogram + hottest SASS lines (stall samples) per kernel from `ncu --page source --csv --print-source sass`.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass | python scripts/ncu_sass_hist.py [top_n]"""
import csv, sys, collections
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
kern, hdr = None, None
data = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        kern = r[1][:70]; data[kern] = []; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if kern and hdr and len(r) == len(hdr):
        data[kern].append(dict(zip(hdr, r)))
for kern, lines in data.items():
    print("====", kern)
    ops, samp = collections.Counter(), collections.Counter()
    tot_i = tot_s = 0
    for d in lines:
        src = d["Source"].strip()
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = op.split(".")[0] if not op.startswith("MUFU") else op
        n = int(d["Instructions Executed"] or 0); s = int(d["# Samples"] or 0)
        ops[op] += n; samp[op] += s; tot_i += n; tot_s += s
    print(f"  warp instructions executed {tot_i}, stall samples {tot_s}")
    for op, n in ops.most_common(top):
        print(f"    {op:14s} {n:12d} {100.0 * n / max(tot_i, 1):5.1f}%   samples {100.0 * samp[op] / max(tot_s, 1):5.1f}%")
    print("  -- hottest lines by stall samples")
    for d in sorted(lines, key=lambda d: -int(d["# Samples"] or 0))[:top]:
        print(f"    {int(d['# Samples']):6d} {100.0 * int(d['# Samples']) / max(tot_s, 1):5.1f}%  exec {int(d['Instructions Executed'] or 0):9d}  {d['Source'].strip()[:90]}")
