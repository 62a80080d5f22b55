set -u
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mf_cs_rows -c 4 -o $O/r02h_ncu_rows python tools/kernel_bench.py --grid 256 --box 32 --reps 2 --only "ghosts from own" > $O/r02h_ncu.log 2>&1
