#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_n2_c.log 2>&1
timeout 600 $TR --master-port 29551 tools/amr_dist_check.py > $O/amr_dist_check_c.log 2>&1
timeout 600 $TR --master-port 29552 tools/amr_dist_check.py --subcycle > $O/amr_dist_check_subcycle_c.log 2>&1
timeout 600 $TR --master-port 29553 tools/amr_bench.py --grid 256 --levels 2 --steps 16 --coupling subcycle > $O/amr_n2_sub_c.json 2> $O/amr_n2_c.err
timeout 600 $TR --master-port 29554 tools/amr_bench.py --grid 256 --levels 3 --steps 12 --regrid-every 4 > $O/amr_n2_3l_regrid_c.json 2>> $O/amr_n2_c.err
timeout 600 $TR --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2_c.json 2>> $O/amr_n2_c.err
