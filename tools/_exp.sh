O=gpurun_out/exp7; mkdir -p $O
run() { tag=$1; port=$2; shift 2; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --no-cpu-baseline "$@" > $O/$tag.json 2> $O/$tag.err; }
run weak256_slabx_xyz 29561 --decomp slab-x --axis-order xyz
run weak256_slabx_zyx 29562 --decomp slab-x --axis-order zyx
run strong512_slabx_xyz 29563 --decomp slab-x --axis-order xyz --size 512 --scaling strong --steps 60
for f in $O/*.json; do echo $f; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"exposed_frac": [0-9.]*\|"axis_order": "[a-z]*"\|"timeline_ms": {[^}]*}' $f | tr '\n' ' '; echo; done
