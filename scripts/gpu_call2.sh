#!/bin/bash
# round-2 GPU visit (2 GPUs): validate the fp16-only peer tail (SNB_PEER_F16ONLY=1) and time the full schedule with both variants
set -u
mkdir -p gpurun_out
make -s -C oracle
bash scripts/gpu_experiments.sh dp
for f in 0 1; do
  SNB_PEER_F16ONLY=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$f \
      scripts/time_to_mesh.py > gpurun_out/ttm_n2_f16only$f.json 2> gpurun_out/ttm_n2_f16only$f.err
  echo "ttm SNB_PEER_F16ONLY=$f rc=$?"; grep "^{" gpurun_out/ttm_n2_f16only$f.json | cut -c1-600
done
tail -5 gpurun_out/*.err | cut -c1-400
