#!/bin/bash
# round-2 first GPU visit: probe for tiny-cuda-nn, L2 reduction-rate microbenchmark, validation of the three dark variants
set -u
mkdir -p gpurun_out
echo "== probe"; python - <<'PY' 2>&1 | tee gpurun_out/probe_modules.txt
import importlib
for m in ("tinycudann", "nerfacc", "mcubes", "open3d", "trimesh", "pyhocon", "pyexr"):
    try:
        importlib.import_module(m); print(m, "PRESENT")
    except Exception as e:
        print(m, "absent:", type(e).__name__)
PY
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "== red_rate"; ./scripts/micro/red_rate 2>&1 | tee gpurun_out/red_rate.txt
bash scripts/gpu_experiments.sh 2>&1 | tee gpurun_out/experiments.txt
