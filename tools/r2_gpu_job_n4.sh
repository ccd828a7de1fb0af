#!/bin/bash
# round 2, 4 GPUs: the z-slab overlap cost at N=4 (round 1: 0.954 efficiency), pencils, strong scaling, fp64
O=gpurun_out/r2n4; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 4 --steps 400 --warmup 20 > $O/weak256_slab.json 2> $O/weak256_slab.err
timeout 300 $TR --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > $O/weak256_slab_driver_like.json 2> $O/weak256_slab_driver_like.err
timeout 300 $TR --master-port 29523 bench.py --gpus 4 --steps 400 --warmup 20 --decomp block > $O/weak256_pencil.json 2> $O/weak256_pencil.err
timeout 300 $TR --master-port 29524 bench.py --gpus 4 --steps 400 --warmup 20 --decomp slab-x > $O/weak256_slabx.json 2> $O/weak256_slabx.err
timeout 300 $TR --master-port 29525 bench.py --gpus 4 --steps 300 --warmup 10 --config 2 > $O/strong512_slab.json 2> $O/strong512_slab.err
timeout 300 $TR --master-port 29526 bench.py --gpus 4 --steps 300 --warmup 10 --config 4 > $O/f64_384_slab.json 2> $O/f64_384_slab.err
timeout 300 $TR --master-port 29527 bench.py --gpus 4 --steps 400 --warmup 20 --graph > $O/weak256_slab_graph.json 2> $O/weak256_slab_graph.err
python tools/summarise_bench.py $O/*.json
