#!/bin/bash
# One parameterised GPU session (replaces the round-1 gpu_session_*.sh one-shots).  Run under gpurun:
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh tests bench shapes kernels'
# Each word selects a stage; outputs land in gpurun_out/<tag>_*.  TAG (env) names the files (default r02).
set -u
O=gpurun_out
T=${TAG:-r02}
mkdir -p $O
for stage in "$@"; do
  case $stage in
    tests)   ( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/${T}_pytest_gpu.log 2>&1 ;;
    smoke)   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1 ;;
    bench)   timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err ;;
    benchref) timeout 900 python bench.py --impl reference > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err ;;
    shapes)  timeout 600 python tools/shape_bench.py --dims 256 256 256 --dims 1024 1024 128 --dims 1024 1024 64 \
               --dims 512 512 512 --dims 1024 256 512 > $O/${T}_shapes.jsonl 2> $O/${T}_shapes.err
             timeout 600 python tools/shape_bench.py --scheme slab --dims 256 256 256 --dims 1024 1024 128 \
               >> $O/${T}_shapes.jsonl 2>> $O/${T}_shapes.err ;;
    kernels) for b in 32 64; do timeout 600 python tools/kernel_bench.py --grid 256 --box $b ; done \
               > $O/${T}_kernel_roofline.jsonl 2> $O/${T}_kernels.err ;;
    amr)     : > $O/${T}_amr.jsonl
             timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 20 2>/dev/null | grep '^{' >> $O/${T}_amr.jsonl
             timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 16 --coupling subcycle 2>/dev/null | grep '^{' >> $O/${T}_amr.jsonl
             timeout 300 python tools/amr_bench.py --grid 256 --levels 3 --steps 12 --regrid-every 4 2>/dev/null | grep '^{' >> $O/${T}_amr.jsonl ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
               --log-file $O/${T}_launches_bench.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-amr > $O/${T}_launches_bench.out 2>&1 ;;
    dist)    NG=${NGPUS:-2}
             ( time timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_slab.py -x -q -m gpu ) > $O/${T}_pytest_dist.log 2>&1 ;;
    benchN)  NG=${NGPUS:-2}
             timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 \
               bench.py --gpus $NG --steps ${STEPS:-20} --warmup 3 > $O/${T}_bench_n$NG.json 2> $O/${T}_bench_n$NG.err ;;
    *) echo "unknown stage $stage" ;;
  esac
done
