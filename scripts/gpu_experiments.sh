#!/bin/bash
# One GPU-box visit that validates and times the experimental variants written at the end of round 1 without GPU access
# (DESIGN.md section 8).  Each is OFF by default; flip the default only after this passes.
#   1 GPU :  bash scripts/gpu_experiments.sh
#   2 GPUs:  gpurun --gpus 2 -- 'bash scripts/gpu_experiments.sh dp'
set -u
mkdir -p gpurun_out
make -s -C oracle
if [ "${1:-}" = "dp" ]; then
  for f in 0 1; do
    SNB_PEER_F16ONLY=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$f \
        scripts/dp_peer_check.py --every 20 > gpurun_out/dp_peer_check_f16only$f.json 2> gpurun_out/dp_peer_check_f16only$f.err
    echo "SNB_PEER_F16ONLY=$f rc=$?"; grep "^{" gpurun_out/dp_peer_check_f16only$f.json
  done
  exit 0
fi
echo "== opt-in tests"; SNB_EXPERIMENTAL=1 python -m pytest tests -m gpu -q -k "peer_tail_world1" 2>&1 | grep -E "passed|failed|^FAILED|^E  " | head
echo "== opt-in tests"; SNB_EXPERIMENTAL=1 python -m pytest tests -m gpu -q -k peer_tail_world1 2>&1 | grep -E "passed|failed|^FAILED|^E  " | head
echo "== baseline"; python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED" | head
echo "== SNB_RENDER_FAST=1"; SNB_RENDER_FAST=1 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED" | head
echo "== SNB_OCC_MMA=1"; SNB_OCC_MMA=1 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED" | head
echo "== step times: baseline, then both switches"
python scripts/kernel_times.py 100 1000 4800 | cut -c1-420
SNB_RENDER_FAST=1 SNB_OCC_MMA=1 python scripts/kernel_times.py 100 1000 4800 | cut -c1-420
