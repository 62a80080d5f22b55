#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_amrsim.py -x -q -k "fused_rohde" 2>&1 | tail -5 > $O/tiling_test.log
: > $O/tiling.jsonl
for g in 128 256; do
 for t in rows linear; do
  for sk in 0 1 2; do
    timeout 300 python tools/amr_bench.py --grid $g --levels 2 --steps 20 --valid-tiling $t --debug-skip $sk 2>/dev/null | grep '^{' >> $O/tiling.jsonl
  done
 done
done
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 20 --valid-tiling linear --max-grid 64 2>/dev/null | grep '^{' >> $O/tiling.jsonl
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 20 --valid-tiling rows --max-grid 64 2>/dev/null | grep '^{' >> $O/tiling.jsonl
