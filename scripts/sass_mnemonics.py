"""SASS mnemonic counts per kernel of libsnb200.so -> profiles/r01_sass_mnemonics.txt style report.
usage: python scripts/sass_mnemonics.py [out.txt]   (cuobjdump -sass; runs without a GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "supernormal_b200", "lib", "libsnb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, stats = None, collections.defaultdict(collections.Counter)
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    m = pat.match(l)
    if m and cur:
        stats[cur][m.group(1)] += 1


def dem(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:110]
    except Exception:
        return n


lines = ["SASS mnemonic counts per kernel of libsnb200.so (cuobjdump -sass, sm_100a): tensor-core (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,",
         "HMMA = mma.sync), MUFU, reductions/atomics, system-scope / strong memory operations (peer-memory kernel).", ""]
for fn, c in sorted(stats.items(), key=lambda kv: dem(kv[0])):
    row = {}
    for op, n in c.items():
        if op.startswith(("UTC", "LDTM", "STTM", "HMMA", "MUFU", "RED", "ATOM", "MEMBAR", "SYNCS", "CCTL")) or ".SYS" in op or "STRONG" in op:
            row[op] = row.get(op, 0) + n
    lines.append(dem(fn))
    lines.append(f"    instructions {sum(c.values())}; " + ", ".join(f"{k} {v}" for k, v in sorted(row.items())))
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_sass_mnemonics.txt")
open(out, "w").write("\n".join(lines) + "\n")
print(out, len(stats), "kernels")
