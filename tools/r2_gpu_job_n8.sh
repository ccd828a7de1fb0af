#!/bin/bash
# round 2, 8 GPUs: 8-rank parity tests, C++ 8-process run, and every BASELINE multi-GPU configuration re-measured
O=gpurun_out/r2n8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multiprocess.py -m gpu -q -x -k "eight_ranks" > $O/pytest_8gpu.log 2>&1; echo "rc=$?" >> $O/pytest_8gpu.log
timeout 300 python -m pytest tests/test_host_cpp.py -m gpu -q -x -k "ipc_ranks" > $O/pytest_ipc_8gpu.log 2>&1; echo "rc=$?" >> $O/pytest_ipc_8gpu.log
B="timeout 300 $TR"
$B --master-port 29531 bench.py --gpus 8 --steps 400 --warmup 20 > $O/weak256_slab.json 2> $O/weak256_slab.err
$B --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 > $O/weak256_slab_driver_like.json 2> $O/weak256_slab_driver_like.err
$B --master-port 29533 bench.py --gpus 8 --steps 200 --warmup 10 --config 3-slab > $O/weak512_slab.json 2> $O/weak512_slab.err
$B --master-port 29534 bench.py --gpus 8 --steps 200 --warmup 10 --config 3-block > $O/weak512_block.json 2> $O/weak512_block.err
$B --master-port 29535 bench.py --gpus 8 --steps 200 --warmup 10 --config 3-pencil > $O/weak512_pencil.json 2> $O/weak512_pencil.err
$B --master-port 29536 bench.py --gpus 8 --steps 400 --warmup 20 --config 2 > $O/strong512_slab.json 2> $O/strong512_slab.err
$B --master-port 29537 bench.py --gpus 8 --steps 400 --warmup 20 --config 4 > $O/f64_384_slab.json 2> $O/f64_384_slab.err
$B --master-port 29538 bench.py --gpus 8 --steps 400 --warmup 20 --decomp block > $O/weak256_block.json 2> $O/weak256_block.err
$B --master-port 29539 bench.py --gpus 8 --steps 200 --warmup 10 --config recipe-weak > $O/recipe_weak.json 2> $O/recipe_weak.err
$B --master-port 29540 bench.py --gpus 8 --steps 400 --warmup 20 --config 2 --graph > $O/strong512_slab_graph.json 2> $O/strong512_slab_graph.err
$B --master-port 29541 bench.py --gpus 8 --min-seconds 5 > $O/weak256_slab_sustained.json 2> $O/weak256_slab_sustained.err
tail -3 $O/pytest_8gpu.log; tail -3 $O/pytest_ipc_8gpu.log
python tools/summarise_bench.py $O/*.json
