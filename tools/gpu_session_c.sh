#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu_c.log 2>&1
timeout 300 python tools/align_test.py > $O/align_test_after.log 2>&1
timeout 300 python tools/kernel_bench.py --grid 256 --box 32 > $O/kernel_bench_256_b32_aligned.jsonl 2> $O/kernel_bench.err
: > $O/amr_c.jsonl
timeout 300 python tools/amr_bench.py --grid 128 --levels 2 --steps 20 2>/dev/null | grep '^{' >> $O/amr_c.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 20 2>/dev/null | grep '^{' >> $O/amr_c.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 3 --steps 10 2>/dev/null | grep '^{' >> $O/amr_c.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 3 --steps 12 --regrid-every 4 2>/dev/null | grep '^{' >> $O/amr_c.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 16 --coupling subcycle 2>/dev/null | grep '^{' >> $O/amr_c.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 16 --coupling subcycle --no-fusion 2>/dev/null | grep '^{' >> $O/amr_c.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 3 --steps 8 --coupling subcycle 2>/dev/null | grep '^{' >> $O/amr_c.jsonl
timeout 300 python bench.py --no-cpu > $O/bench_c.json 2>/dev/null
