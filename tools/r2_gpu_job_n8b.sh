#!/bin/bash
# round 2, 8 GPUs, after the lane-code rework: 8-rank parity, x-cutting decompositions re-measured
O=gpurun_out/r2n8b; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -m gpu -q -x -k "eight_ranks" > $O/pytest_8gpu.log 2>&1; echo "rc=$?" >> $O/pytest_8gpu.log
B="timeout 300 $TR"
$B --master-port 29538 bench.py --gpus 8 --steps 400 --warmup 20 --decomp block > $O/weak256_block.json 2> $O/weak256_block.err
$B --master-port 29534 bench.py --gpus 8 --steps 200 --warmup 10 --config 3-block > $O/weak512_block.json 2> $O/weak512_block.err
$B --master-port 29539 bench.py --gpus 8 --steps 200 --warmup 10 --config recipe-weak > $O/recipe_weak.json 2> $O/recipe_weak.err
$B --master-port 29540 bench.py --gpus 8 --steps 400 --warmup 20 --decomp slab-x > $O/weak256_slabx.json 2> $O/weak256_slabx.err
$B --master-port 29531 bench.py --gpus 8 --steps 400 --warmup 20 > $O/weak256_slab.json 2> $O/weak256_slab.err
tail -3 $O/pytest_8gpu.log
python tools/summarise_bench.py $O/*.json
