#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/amr_dist_check.py --subcycle > $O/amr_dist_check_subcycle.log 2>&1
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 16 --coupling subcycle > $O/amr_sub_2l_256.json 2> $O/amr_sub.err
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 16 --coupling subcycle --gradient 2e-4 --regrid-every 8 > $O/amr_sub_2l_256_grad.json 2>> $O/amr_sub.err
timeout 300 python tools/amr_bench.py --grid 256 --levels 3 --steps 8 --coupling subcycle > $O/amr_sub_3l_256.json 2>> $O/amr_sub.err
timeout 600 $TR --master-port 29522 tools/amr_bench.py --grid 256 --levels 2 --steps 16 --coupling subcycle --gradient 2e-4 --regrid-every 8 > $O/amr_sub_n2_2l_256_grad.json 2>> $O/amr_sub.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_amr_sub_2l_128.csv \
  python tools/amr_bench.py --grid 128 --levels 2 --steps 4 --warmup 3 --coupling subcycle > $O/ncu_amr_sub.log 2>&1
