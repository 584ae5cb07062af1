#!/bin/bash
# the driver's scaling command at N GPUs (bench only)
set -u
N=${1:-8}
mkdir -p gpurun_out
make -s -C oracle
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu \
    > gpurun_out/bench_n${N}_w5k20.json 2> gpurun_out/bench_n${N}_w5k20.err
echo "bench N=$N rc=$?"; python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_n${N}_w5k20.json"))
    for k in ("value", "ms_per_step", "e2e", "continuation", "time_to_mesh", "schedule_avg", "clocks", "gpu_launches"):
        print(k, json.dumps(j.get(k))[:400])
except Exception as e:
    print("no json", e)
PY
tail -c 600 gpurun_out/bench_n${N}_w5k20.err
