#!/bin/bash
set -u
mkdir -p gpurun_out
make -s -C oracle
python -m pytest tests -m gpu -x -q -k "fused_forward_backward or fma_backward or training_reduces or device_sampler" 2>&1 | tail -8
echo "== kernel times (split backward v2)"; python scripts/kernel_times.py 1500 2500 4800 | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_it4800.csv python scripts/ncu_at.py 4800 3 > gpurun_out/ncu_l.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_it4800.csv | tee gpurun_out/launches_it4800_summary.txt
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'sdf_bwd_mlp_umma|hash_scatter' -c 2 -o gpurun_out/prof_bwd_split_v2_it4800 python scripts/ncu_at.py 4800 1 > gpurun_out/ncu_f.log 2>&1
ncu -i gpurun_out/prof_bwd_split_v2_it4800.ncu-rep --page raw --csv 2>/dev/null | python scripts/summarize_ncu_raw.py | tee gpurun_out/prof_bwd_split_v2_it4800.txt
