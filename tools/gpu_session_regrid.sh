#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_b.log 2>&1
LBX_HOST_TIMING=1 timeout 300 python tools/amr_bench.py --grid 256 --levels 3 --steps 8 --regrid-every 4 > $O/regrid_prof_256.json 2> $O/regrid_prof_256.err
LBX_HOST_TIMING=1 timeout 300 python tools/amr_bench.py --grid 384 --levels 3 --steps 4 --regrid-every 2 > $O/regrid_prof_384.json 2> $O/regrid_prof_384.err
timeout 120 lambrex_b200/_lib/derived_vars 32 50 > $O/derived_vars.log 2>&1
