#!/bin/bash
# BASELINE configs[2]: own_objects.conf-shaped data-parallel training on N GPUs (36 views 2016x1512, 30 000-iteration schedule)
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 200 --warmup 50 --no-cpu \
    --workload own_objects > gpurun_out/bench_own_objects_n${N}.json 2> gpurun_out/bench_own_objects_n${N}.err
echo "rc=$?"; python - <<PY
import json
j = json.load(open("gpurun_out/bench_own_objects_n${N}.json"))
for k in ("value", "ms_per_step", "e2e", "continuation", "time_to_mesh", "schedule_avg", "clocks"):
    print(k, json.dumps(j.get(k))[:400])
PY
tail -c 800 gpurun_out/bench_own_objects_n${N}.err
