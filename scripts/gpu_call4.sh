#!/bin/bash
set -u
mkdir -p gpurun_out
make -s -C oracle
python -m pytest tests -m gpu -x -q -k "fused_forward_backward or smoke or fma_backward or training_reduces" 2>&1 | tail -15
echo "== kernel times (split backward)"; python scripts/kernel_times.py 1500 2500 4800 | cut -c1-600
echo "== kernel times (fused backward)"; SNB_BWD_SPLIT=0 python scripts/kernel_times.py 1500 2500 4800 | cut -c1-600
