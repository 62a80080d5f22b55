#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29571 tools/amr_dist_check.py > $O/amr_dist_check_d.log 2>&1
timeout 300 $TR --master-port 29572 tools/amr_dist_check.py --subcycle > $O/amr_dist_check_subcycle_d.log 2>&1
timeout 300 $TR --master-port 29573 tools/amr_bench.py --grid 256 --levels 3 --steps 12 --regrid-every 4 > $O/amr_n2_3l_regrid_d.json 2> $O/amr_n2_d.err
timeout 300 $TR --master-port 29574 tools/amr_bench.py --grid 256 --levels 2 --steps 20 > $O/amr_n2_2l_d.json 2>> $O/amr_n2_d.err
