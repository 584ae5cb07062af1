#!/bin/bash
# compute-sanitizer (memcheck, racecheck) on smoke() + ncu --set full of the data-parallel tail kernel with a one-rank peer group
set -u
mkdir -p gpurun_out
make -s -C oracle
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|smoke ok|Invalid|out of bounds" gpurun_out/sanitizer_memcheck.log | head
timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|smoke ok|hazard" gpurun_out/sanitizer_racecheck.log | head
timeout 300 ncu --set full --clock-control none -k regex:train_tail_peer -c 3 -o gpurun_out/prof_tail_peer_world1 python -m pytest tests/test_gpu_fused.py -q -k "peer_tail_world1" > gpurun_out/ncu_tail_peer.log 2>&1
ncu -i gpurun_out/prof_tail_peer_world1.ncu-rep --page raw --csv 2>/dev/null | python scripts/summarize_ncu_raw.py | tee gpurun_out/prof_tail_peer_world1.txt | grep -E "^----|duration|dram__bytes|issue_active"
