#!/bin/bash
# Eight-GPU session (gpurun --gpus 8): BASELINE configs[2] at N=8 and configs[4] (3-level AMR, 512^3 base,
# periodic regrid, boxes distributed over 8 GPUs) + the distributed bit-equality check.
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo8.txt 2>&1
timeout 600 $TR --master-port 29531 bench.py --gpus 8 --steps 30 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
timeout 600 $TR --master-port 29532 tools/amr_bench.py --grid 512 --levels 3 --steps 32 --regrid-every 16 > $O/amr_n8_3l_512_regrid.json 2> $O/amr_n8.err
timeout 400 $TR --master-port 29533 tools/amr_dist_check.py > $O/amr_dist_check_n8.log 2>&1
timeout 400 $TR --master-port 29534 tools/amr_bench.py --grid 512 --levels 3 --steps 8 --coupling subcycle --gradient 2e-4 --regrid-every 4 > $O/amr_n8_3l_512_subcycle_grad.json 2>> $O/amr_n8.err
ls -la $O
