for cfg in "12 0" "12 1" "12 2" "12 6" "12 4" "0 0" "0 1" "0 2"; do set -- $cfg; echo "WARPS=$1 DBG=$2"; SNB_BWD_WARPS=$1 SNB_BWD_DBG=$2 python scripts/kernel_times.py 1000 4800 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['iter'], d['us']['snb_sdf_bwd_patch'])"; done
