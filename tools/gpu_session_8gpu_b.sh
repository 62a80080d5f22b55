#!/bin/bash
# Eight-GPU session 2: BASELINE configs[4] again after the regrid work (3-level AMR, 512^3 base, boxes over 8 GPUs)
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29561 tools/amr_bench.py --grid 512 --levels 3 --steps 48 --regrid-every 16 > $O/amr_n8_3l_512_regrid_b.json 2> $O/amr_n8_b.err
timeout 400 $TR --master-port 29562 tools/amr_bench.py --grid 512 --levels 3 --steps 16 > $O/amr_n8_3l_512_static_b.json 2>> $O/amr_n8_b.err
timeout 400 $TR --master-port 29563 tools/amr_bench.py --grid 512 --levels 3 --steps 8 --coupling subcycle > $O/amr_n8_3l_512_subcycle_b.json 2>> $O/amr_n8_b.err
