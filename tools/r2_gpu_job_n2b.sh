#!/bin/bash
# round 2, 2 GPUs, after the lane-code rework: cross-device parity tests, x-cutting bench lines
O=gpurun_out/r2n2b; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multiprocess.py tests/test_gpu_multidomain.py -m gpu -q -x -k "torchrun or across_devices or two_processes or graph" > $O/pytest_2gpu.log 2>&1; echo "rc=$?" >> $O/pytest_2gpu.log
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 400 --warmup 20 --decomp slab-x > $O/weak256_slabx.json 2> $O/weak256_slabx.err
timeout 300 $TR --master-port 29515 bench.py --gpus 2 --steps 200 --warmup 10 --config 2 --decomp slab-x > $O/strong512_slabx.json 2> $O/strong512_slabx.err
timeout 300 $TR --master-port 29519 bench.py --gpus 2 --steps 200 --warmup 10 --config recipe-weak > $O/recipe_weak.json 2> $O/recipe_weak.err
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 400 --warmup 20 > $O/weak256_slab.json 2> $O/weak256_slab.err
tail -3 $O/pytest_2gpu.log
python tools/summarise_bench.py $O/*.json
tail -2 $O/*.err | cut -c1-200
