#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_amr_kernels.py -x -q 2>&1 | tail -3
python tools/kernel_bench.py --only FillBoundary 2>&1 | tail -1
python tools/kernel_bench.py --only FillBoundary --grid 512 --box 64 2>&1 | tail -1
python -m pytest tests/test_gpu_amrsim.py -x -q 2>&1 | tail -3
