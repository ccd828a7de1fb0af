#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
LBM_B200_XTILE_W=0 timeout 120 python tools/probe_timeline.py x > $O/timeline_x_linear.txt 2>&1
LBM_B200_XTILE_W=0 LBM_B200_XFUSE_DEBUG=none timeout 120 python tools/probe_timeline.py x > $O/timeline_x_linear_none.txt 2>&1
LBM_B200_XTILE_W=0 timeout 120 python tools/probe_timeline.py z > $O/timeline_z.txt 2>&1
LBM_B200_XTILE_W=0 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"lbm_(alpha|beta)_kernel|halo" -s 20 -c 12 --csv --log-file $O/ncu_small.csv \
   python tools/probe_overlap.py --axes x --only overlap --steps 8 > $O/ncu_small.log 2>&1
cat $O/timeline_x_linear.txt | head -40; cat $O/ncu_small.csv | tail -30 | cut -c1-200
