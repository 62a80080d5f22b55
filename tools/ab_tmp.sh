set -u
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mf_cs_rows -c 2 -o $O/r02m_ncu_rows python tools/kernel_bench.py --grid 256 --box 32 --reps 1 --only "ghosts from own" > $O/r02m_ncu.log 2>&1
timeout 300 python tools/kernel_bench.py --grid 256 --box 32 --only k_mf_collide_stream > $O/r02m_kernel.jsonl 2>&1
