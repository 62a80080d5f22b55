#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
for lib in _lib _lib_prev; do
echo $lib
LBX_LIB_DIR=$lib python tools/amr_bench.py --grid 256 --levels 2 --steps 20 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('MLUPS %.0f ms/step %.3f kernel frac %.3f share %.2f' % (d['value'], d['ms_per_step'], r['frac'], r['share_of_timed_region']))"
LBX_LIB_DIR=$lib python tools/kernel_bench.py --only "plan_apply" --row-kernel 0 2>&1 | grep -o '"kernel": "[^"]*\|"ms": [0-9.]*\|"frac[^,]*' | paste - - -
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_amr_2level_256.csv python tools/amr_bench.py --grid 256 --levels 2 --steps 6 --warmup 3 > $O/l1.out 2>&1
