#!/bin/bash
# One-GPU measurement session (run under gpurun): GPU tests, bench lines, AMR-path lines,
# ncu launch lists and one --set full capture per hot kernel.  Outputs -> gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $O/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python bench.py --api raw --scheme pull --no-cpu > $O/bench_raw_pull.json 2>> $O/bench_n1.err
timeout 300 python bench.py --api raw --scheme push --no-cpu > $O/bench_raw_push.json 2>> $O/bench_n1.err
timeout 300 python bench.py --grid 512 --no-cpu --steps 50 > $O/bench_512.json 2>> $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_ref.json 2>> $O/bench_n1.err
# AMR path (configs[3], configs[4] on one GPU)
timeout 300 python tools/amr_bench.py --grid 128 --levels 2 --steps 20 > $O/amr_2l_128.json 2> $O/amr.err
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 20 > $O/amr_2l_256.json 2>> $O/amr.err
timeout 300 python tools/amr_bench.py --grid 128 --levels 3 --steps 10 > $O/amr_3l_128.json 2>> $O/amr.err
timeout 300 python tools/amr_bench.py --grid 256 --levels 3 --steps 10 --regrid-every 4 > $O/amr_3l_256_regrid.json 2>> $O/amr.err
timeout 300 python tools/amr_bench.py --grid 256 --levels 2 --steps 10 --no-fusion > $O/amr_2l_256_nofusion.json 2>> $O/amr.err
# launch lists
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_256.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu --e2e-repeat 1 > $O/ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_amr_2l_128.csv \
  python tools/amr_bench.py --grid 128 --levels 2 --steps 4 --warmup 3 > $O/ncu_amr.log 2>&1
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide_stream -s 5 -c 1 -o $O/full_collide_stream \
  python bench.py --steps 10 --warmup 3 --no-cpu --e2e-repeat 1 > $O/ncu_full_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mf_collide_stream -s 12 -c 2 -o $O/full_mf_collide_stream \
  python tools/amr_bench.py --grid 128 --levels 2 --steps 4 --warmup 3 > $O/ncu_full_amr.log 2>&1
ls -la $O
