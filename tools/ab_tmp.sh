#!/bin/bash
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
( time timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_slab.py -x -q -m gpu ) 2>&1 | tail -4
timeout 400 $TR tools/amr_dist_check.py 2>&1 | grep -v "INITIAL GRIDS" | tail -12
$TR tools/amr_bench.py --grid 256 --levels 3 --steps 12 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=2 3-level MLUPS %.0f ms/step %.3f kernel frac %.3f share %.2f' % (d['value'], d['ms_per_step'], r['frac'], r['share_of_timed_region']))"
