#!/bin/bash
set -u
mkdir -p gpurun_out
make -s -C oracle
python -m pytest tests -m gpu -q -x -k "mesh_post_on_device or ad_gradient" 2>&1 | grep -E "^E  |^>|passed|failed|Error" | head -40
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_w5k20.json 2> gpurun_out/bench_w5k20.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_w5k20.err; python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_w5k20.json"))
for k in ("value", "ms_per_step", "e2e", "continuation", "time_to_mesh", "schedule_avg", "roofline", "roofline_l2_reduction", "cpu_baseline", "gpu_launches", "clocks", "dtype"):
    print(k, json.dumps(j.get(k))[:600])
print("ref", json.dumps(j.get("reference_cuda_path"))[:300])
PY
bash scripts/gpu_ncu_all.sh 4800 | head -80
