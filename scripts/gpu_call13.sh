#!/bin/bash
set -u
mkdir -p gpurun_out
make -s -C oracle
python -m pytest tests -m gpu -q 2>&1 | tail -5
bash scripts/capture_roofline_traffic.sh
python scripts/kernel_times.py 100 4800 | cut -c1-400
