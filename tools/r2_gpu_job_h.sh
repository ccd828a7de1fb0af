#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lbm_(alpha|beta)_kernel" -s 12 -c 2 -o $O/xfuse_x256 \
   python tools/probe_overlap.py --axes x --only overlap --steps 8 > $O/ncu_xfuse.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lbm_(alpha|beta)_kernel" -s 12 -c 2 -o $O/plain_256 \
   python tools/probe_overlap.py --axes x --only plain --steps 8 > $O/ncu_plain.log 2>&1
ls -la $O
