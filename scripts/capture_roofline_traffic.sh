#!/bin/bash
# dram__bytes_read + dram__bytes_write per launch of the dominant kernel ON THE WINDOW THE DRIVER TIMES (bench.py --steps 20 --warmup 5):
# one `ncu --set full` capture of its launches inside the timed region (bench.py brackets it with cudaProfilerStart/Stop).
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'sdf_bwd_patch_umma_kernel' -c 6 -o gpurun_out/prof_roofline_traffic \
    python bench.py --steps 20 --warmup 5 --no-cpu --no-ttm --e2e-steps 1 --ref-cuda-steps 0 --cont-steps 0 > gpurun_out/ncu_traffic.log 2>&1
ncu -i gpurun_out/prof_roofline_traffic.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/prof_roofline_traffic_raw.csv
python scripts/summarize_ncu_raw.py < gpurun_out/prof_roofline_traffic_raw.csv > gpurun_out/prof_roofline_traffic.txt
python - <<'PY'
import csv, json
rows = list(csv.reader(l for l in open("gpurun_out/prof_roofline_traffic_raw.csv") if not l.startswith("==")))
hdr, units = rows[0], rows[1]
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
tot, n, names = 0.0, 0, set()
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    tot += to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
    n += 1
    names.add(d["Kernel Name"].split("(")[0])
out = {"entry_point": "snb_sdf_bwd_patch_ws", "kernel": sorted(names), "launches": n, "dram_bytes_per_launch": tot / max(n, 1),
       "command": "ncu --set full --clock-control none --profile-from-start off -k regex:sdf_bwd_patch_umma_kernel -c 6 python bench.py --steps 20 --warmup 5 (timed region only)",
       "note": f"mean of {n} launches inside the driver's window (iterations 5-25, 1 live level), ncu --set full --clock-control none; inputs written by the "
               "forward / render kernels of the same step are mostly still L2-resident, so this can sit below the algorithmic bytes"}
json.dump(out, open("gpurun_out/roofline_traffic.json", "w"), indent=1)
print(json.dumps(out))
PY
