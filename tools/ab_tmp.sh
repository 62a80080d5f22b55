set -u
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r03b_pytest_gpu.log 2>&1
timeout 300 python tools/shape_bench.py --scheme push --dims 256 256 256 > $O/r03b_shape.jsonl 2>&1
