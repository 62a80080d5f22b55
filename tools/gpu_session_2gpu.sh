#!/bin/bash
# Two-GPU session (gpurun --gpus 2): slab-path and distributed-AMR checks + bench lines.
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_n2.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu_n2.log
timeout 600 $TR --master-port 29511 tools/amr_dist_check.py > $O/amr_dist_check.log 2>&1
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 3 --halo nccl > $O/bench_n2_nccl.json 2>> $O/bench_n2.err
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/bench_n2_ref.json 2>> $O/bench_n2.err
timeout 600 $TR --master-port 29515 tools/amr_bench.py --grid 256 --levels 2 --steps 20 > $O/amr_n2_2l_256.json 2> $O/amr_n2.err
timeout 600 $TR --master-port 29516 tools/amr_bench.py --grid 256 --levels 3 --steps 12 --regrid-every 4 > $O/amr_n2_3l_256_regrid.json 2>> $O/amr_n2.err
ls -la $O
