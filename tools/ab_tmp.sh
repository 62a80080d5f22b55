#!/bin/bash
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655"
show() { grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print(d['n_gpus'], d['levels'], d['coupling'], 'MLUPS %.0f ms/step %.3f kernel frac %.3f share %.2f launches/step %.1f' % (d['value'], d['ms_per_step'], r['frac'], r['share_of_timed_region'], d['launches_per_step']))
"; }
( time timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_slab.py tests/test_gpu_amr_kernels.py tests/test_gpu_amrsim.py -x -q -m gpu ) 2>&1 | tail -6
echo "new N=2 3-level 256"; $TR tools/amr_bench.py --grid 256 --levels 3 --steps 12 2>/dev/null | show
echo "prev N=2 3-level 256"; LBX_LIB_DIR=_lib_prev $TR tools/amr_bench.py --grid 256 --levels 3 --steps 12 2>/dev/null | show
echo "new N=1 3-level 256"; python tools/amr_bench.py --grid 256 --levels 3 --steps 12 2>/dev/null | show
echo "new N=2 2-level 256 subcycle"; $TR tools/amr_bench.py --grid 256 --levels 2 --steps 12 --coupling subcycle 2>/dev/null | show
echo "prev N=2 2-level 256 subcycle"; LBX_LIB_DIR=_lib_prev $TR tools/amr_bench.py --grid 256 --levels 2 --steps 12 --coupling subcycle 2>/dev/null | show
timeout 300 $TR tools/amr_dist_check.py 2>&1 | tail -5
