#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
TAG=r02f NGPUS=8 STEPS=20 bash tools/gpu_session.sh benchN
tail -c 600 $O/r02f_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29677 tools/amr_dist_check.py > $O/r02f_amr_dist_check_n8.log 2>&1
grep -v "INITIAL GRIDS\|NX:\|NZ:\|^$" $O/r02f_amr_dist_check_n8.log | tail -8
