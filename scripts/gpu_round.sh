#!/bin/bash
# One GPU-box visit in the driver's order: parity tests, smoke, the reference arm and our arm of bench.py as the driver launches them.
set -u
mkdir -p gpurun_out
make -s -C oracle
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps ${STEPS:-20} --warmup ${WARMUP:-5} > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/bench_reference.json
python bench.py --gpus 1 --steps ${STEPS:-20} --warmup ${WARMUP:-5} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "our arm rc=$?"
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench.json"))
for k in ("value", "ms_per_step", "e2e", "continuation", "time_to_mesh", "schedule_avg", "roofline", "roofline_l2_reduction", "cpu_baseline", "gpu_launches", "clocks", "dtype"):
    print(k, json.dumps(j.get(k))[:420])
for k in ("reference_cuda_path", "dropin_api_path"):
    r = j.get(k) or {}
    print(k, r.get("ms_per_step"), r.get("ours_ms_per_step_same_window"))
PY
