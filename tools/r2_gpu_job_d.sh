#!/bin/bash
# round 2, GPU job D (1 GPU): where does the XFUSE kernel time go?  ncu plain vs fused
O=gpurun_out/r2d; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lbm_(alpha|beta)_kernel" -s 12 -c 4 -o $O/xfuse_x256 \
   python tools/probe_overlap.py --axes x --only overlap --steps 8 > $O/ncu_xfuse.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lbm_(alpha|beta)_kernel" -s 12 -c 4 -o $O/plain_256 \
   python tools/probe_overlap.py --axes x --only plain --steps 8 > $O/ncu_plain.log 2>&1
for f in xfuse_x256 plain_256; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null
done
ls -la $O; tail -3 $O/ncu_xfuse.log
