set -u
O=gpurun_out; mkdir -p $O
LBX_HOST_TIMING=1 timeout 600 python tools/amr_bench.py --grid 256 --levels 3 --steps 32 --regrid-every 16 > $O/r02z_amr3.jsonl 2> $O/r02z_amr3.err
timeout 600 python tools/amr_bench.py --grid 256 --levels 3 --steps 16 >> $O/r02z_amr3.jsonl 2>> $O/r02z_amr3_static.err
