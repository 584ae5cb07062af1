#!/bin/bash
# full GPU test-suite + per-kernel times at three points of the schedule
set -u
mkdir -p gpurun_out
make -s -C oracle
python -m pytest tests -m gpu -q 2>&1 | tail -6
python scripts/kernel_times.py ${ITERS:-100 1000 4800} | cut -c1-520 | tee gpurun_out/kernel_times.txt
