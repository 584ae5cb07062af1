#!/bin/bash
# ncu --set full of every kernel of the step (one launch each) at a given iteration; summaries under gpurun_out/
set -u
IT=${1:-1000}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'march_visible|compact_one|sdf_fwd_patch|render_fused|sdf_bwd|hash_scatter|train_tail|occgrid_update_kernel' \
    -c 9 -o gpurun_out/prof_step_it$IT python scripts/ncu_at.py $IT 1 > gpurun_out/ncu_step_it$IT.log 2>&1
ncu -i gpurun_out/prof_step_it$IT.ncu-rep --page raw --csv 2>/dev/null | python scripts/summarize_ncu_raw.py > gpurun_out/prof_step_it$IT.txt
grep -E "^----|duration|issue_active|warps_active.avg|dram__bytes|registers" gpurun_out/prof_step_it$IT.txt
