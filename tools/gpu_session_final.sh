#!/bin/bash
# Final 1-GPU verification of the round: GPU tests, smoke, the bench line, AMR lines, launch lists.
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_final.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1
timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_ref_final.json 2>> $O/bench_final.err
: > $O/amr_final.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 20 2>/dev/null | grep '^{' >> $O/amr_final.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 3 --steps 12 --regrid-every 4 2>/dev/null | grep '^{' >> $O/amr_final.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 16 --coupling subcycle 2>/dev/null | grep '^{' >> $O/amr_final.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 16 --coupling subcycle --gradient 2e-4 --regrid-every 8 2>/dev/null | grep '^{' >> $O/amr_final.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_amr_2l_128_final.csv \
  python tools/amr_bench.py --grid 128 --levels 2 --steps 4 --warmup 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_amr_sub_2l_128_final.csv \
  python tools/amr_bench.py --grid 128 --levels 2 --steps 4 --warmup 3 --coupling subcycle > /dev/null 2>&1
