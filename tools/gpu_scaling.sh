#!/bin/bash
# usage: tools/gpu_scaling.sh N [extra bench args]   -- runs bench.py on N GPUs (torchrun for N > 1)
N=$1; shift
if [ "$N" = "1" ]; then exec python bench.py --gpus 1 "$@"; fi
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N "$@"
