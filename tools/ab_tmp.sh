set -u
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_amrsim.py tests/test_gpu_amr_kernels.py -m gpu -x -q ) > $O/r02k_pytest_gpu.log 2>&1
for b in 32 64; do timeout 300 python tools/kernel_bench.py --grid 256 --box $b --only k_mf_collide_stream ; done > $O/r02k_kernel_rows.jsonl 2> $O/r02k_kernel_rows.err
timeout 600 python tools/amr_bench.py --grid 256 --levels 2 --steps 12 2>/dev/null | grep '^{' > $O/r02k_amr.jsonl
timeout 600 python tools/amr_bench.py --grid 256 --levels 2 --steps 12 --coupling subcycle 2>/dev/null | grep '^{' >> $O/r02k_amr.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mf_cs_rows -c 2 -o $O/r02k_ncu_rows python tools/kernel_bench.py --grid 256 --box 32 --reps 1 --only "ghosts from own" > $O/r02k_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02k_launches_amr.csv python tools/amr_bench.py --grid 128 --levels 2 --steps 4 --warmup 3 > /dev/null 2>&1
