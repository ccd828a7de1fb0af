#!/bin/bash
# round 2, 2 GPUs: cross-device parity tests, bench lines (z-slabs, x-slabs, strong 512, fp64), graph replay
O=gpurun_out/r2n2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multiprocess.py tests/test_gpu_multidomain.py -m gpu -q -x -k "torchrun or across_devices or two_processes or graph" > $O/pytest_2gpu.log 2>&1; echo "rc=$?" >> $O/pytest_2gpu.log
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 400 --warmup 20 > $O/weak256_slab.json 2> $O/weak256_slab.err
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 400 --warmup 20 --decomp slab-x > $O/weak256_slabx.json 2> $O/weak256_slabx.err
LBM_B200_XFUSE=0 timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 400 --warmup 20 --decomp slab-x > $O/weak256_slabx_nofuse.json 2> $O/weak256_slabx_nofuse.err
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 200 --warmup 10 --config 2 > $O/strong512_slab.json 2> $O/strong512_slab.err
timeout 300 $TR --master-port 29515 bench.py --gpus 2 --steps 200 --warmup 10 --config 2 --decomp slab-x > $O/strong512_slabx.json 2> $O/strong512_slabx.err
LBM_B200_XFUSE=0 timeout 300 $TR --master-port 29516 bench.py --gpus 2 --steps 200 --warmup 10 --config 2 --decomp slab-x > $O/strong512_slabx_nofuse.json 2> $O/strong512_slabx_nofuse.err
timeout 300 $TR --master-port 29517 bench.py --gpus 2 --steps 400 --warmup 20 --graph > $O/weak256_slab_graph.json 2> $O/weak256_slab_graph.err
timeout 300 $TR --master-port 29518 bench.py --gpus 2 --steps 200 --warmup 10 --config 4 > $O/f64_384_slab.json 2> $O/f64_384_slab.err
timeout 300 $TR --master-port 29519 bench.py --gpus 2 --steps 200 --warmup 10 --config recipe-weak > $O/recipe_weak.json 2> $O/recipe_weak.err
tail -3 $O/pytest_2gpu.log
python tools/summarise_bench.py $O/*.json
tail -3 $O/*.err | cut -c1-200
