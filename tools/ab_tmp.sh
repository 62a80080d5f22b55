set -u
O=gpurun_out; mkdir -p $O
timeout 300 python tools/sustained_bench.py --seconds 45 > $O/r02y_sustained.jsonl 2> $O/r02y_sustained.err
