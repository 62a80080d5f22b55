set -u
O=gpurun_out; mkdir -p $O
timeout 600 python tools/shape_bench.py --scheme slabsync --dims 256 256 256 --dims 1024 1024 128 > $O/r02e_shapes_sync.jsonl 2> $O/r02e_shapes_sync.err
timeout 600 python tools/shape_bench.py --scheme slab --dims 1024 1024 128 >> $O/r02e_shapes_sync.jsonl 2>> $O/r02e_shapes_sync.err
for h in p2p slabsim-p2p; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-amr --halo $h > $O/r02e_bench_n2_$h.json 2> $O/r02e_bench_n2_$h.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-amr --grid-multi 512 > $O/r02e_bench_n2_512.json 2> $O/r02e_bench_n2_512.err
