#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list (+ optional full capture of one kernel).
set -u
mkdir -p gpurun_out
make -s -C oracle
python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^E  |Error|^FAILED" | head -60
python bench.py --steps ${STEPS:-400} --warmup ${WARMUP:-20} ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json
if [ "${NCU:-1}" = "1" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-400} --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 60 --warmup ${NCU_WARMUP:-3} --no-cpu --e2e-steps 1 --ref-cuda-steps 0 > gpurun_out/ncu_bench.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.txt
fi
if [ -n "${NCU_KERNEL:-}" ]; then
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${NCU_KERNEL} -s ${NCU_KSKIP:-20} -c 2 -o gpurun_out/prof_${NCU_KERNEL} \
      python bench.py --steps 60 --warmup ${NCU_WARMUP:-3} --no-cpu --e2e-steps 1 --ref-cuda-steps 0 > gpurun_out/ncu_full.log 2>&1
  ncu -i gpurun_out/prof_${NCU_KERNEL}.ncu-rep --page raw --csv 2>/dev/null | python scripts/summarize_ncu_raw.py | tee gpurun_out/prof_${NCU_KERNEL}.txt
fi
