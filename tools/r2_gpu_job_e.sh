#!/bin/bash
# round 2, GPU job E (1 GPU): tiled XFUSE mapping -- parity, then timing matrix over the tile width
O=gpurun_out/r2e; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
for w in 0 4 8 16; do
  echo "== XTILE_W=$w" >> $O/tile.jsonl
  LBM_B200_XTILE_W=$w timeout 120 python tools/probe_overlap.py --axes x >> $O/tile.jsonl 2>> $O/tile.err
  LBM_B200_XTILE_W=$w LBM_B200_XFUSE_DEBUG=none timeout 120 python tools/probe_overlap.py --axes x >> $O/tile.jsonl 2>> $O/tile.err
done
LBM_B200_XTILE_W=8 timeout 120 python tools/probe_overlap.py --axes xyz --timeline >> $O/tile_xyz.jsonl 2>> $O/tile.err
LBM_B200_XTILE_W=8 timeout 120 python tools/probe_overlap.py --axes xyz --size 512 --steps 40 --timeline >> $O/tile_xyz.jsonl 2>> $O/tile.err
LBM_B200_XTILE_W=4 timeout 120 python tools/probe_overlap.py --axes xyz --size 512 --steps 40 >> $O/tile_xyz.jsonl 2>> $O/tile.err
LBM_B200_XTILE_W=8 timeout 120 python tools/probe_overlap.py --axes x --size 1024x1024x32 --steps 100 >> $O/tile_xyz.jsonl 2>> $O/tile.err
LBM_B200_XTILE_W=8 timeout 120 python tools/probe_overlap.py --axes x --size 384 --dtype f64 --steps 60 >> $O/tile_xyz.jsonl 2>> $O/tile.err
tail -4 $O/pytest_gpu.log; cat $O/tile.jsonl $O/tile_xyz.jsonl | cut -c1-330; tail -3 $O/tile.err
