#!/bin/bash
set -u
mkdir -p gpurun_out
make -s -C oracle
python -m pytest tests -m gpu -q -k "feeder or host_fed or e2e or checkpoint" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu --no-ttm --ref-cuda-steps 0 --cont-steps 0 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', j['value'], 'e2e', j['e2e']['value'], 'ms', j['ms_per_step'])"
# BASELINE configs[3] at T = 2^22 (183 MB fp16 table > L2): DRAM bytes and L2 sectors per launch
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum \
    --clock-control none -k regex:'hashgrid_fwd|sdf_eval' --csv --log-file gpurun_out/microbench_t22_ncu.csv python scripts/microbench_sdf.py --log2-n 24 --reps 1 --tables 22 > gpurun_out/microbench_t22.json 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/microbench_t22_ncu.csv") if not l.startswith("=="))]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
idi = hdr.index("ID")
agg = {}
for r in rows[1:]:
    agg.setdefault((r[idi], r[ki][:40]), {})[r[mi]] = (r[vi], r[ui])
for (i, k), m in agg.items():
    print(i, k, {a: " ".join(b) for a, b in m.items()})
PY
