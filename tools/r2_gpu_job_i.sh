#!/bin/bash
# round 2, job i: lane-only XFUSE code paths (3-D grid, per-side pull blocks), multi-block rim pass
O=gpurun_out/r2i; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 120 python tools/probe_overlap.py --axes x >> $O/p.jsonl 2>> $O/p.err
LBM_B200_XFUSE_DEBUG=none timeout 120 python tools/probe_overlap.py --axes x >> $O/p.jsonl 2>> $O/p.err
timeout 120 python tools/probe_overlap.py --axes xyz >> $O/p.jsonl 2>> $O/p.err
timeout 120 python tools/probe_overlap.py --axes xyz --size 512 --steps 40 >> $O/p.jsonl 2>> $O/p.err
timeout 120 python tools/probe_overlap.py --axes x --size 512 --steps 40 >> $O/p.jsonl 2>> $O/p.err
timeout 120 python tools/probe_overlap.py --axes x --size 1024x1024x32 --steps 100 >> $O/p.jsonl 2>> $O/p.err
timeout 120 python tools/probe_overlap.py --axes x --size 384 --dtype f64 --steps 60 >> $O/p.jsonl 2>> $O/p.err
timeout 120 python tools/probe_overlap.py --axes z >> $O/p.jsonl 2>> $O/p.err
timeout 120 python tools/probe_timeline.py x > $O/timeline_x.txt 2>&1
timeout 120 python tools/probe_timeline.py xyz > $O/timeline_xyz.txt 2>&1
timeout 120 python tools/probe_timeline.py xyz 512x512x512 > $O/timeline_xyz512.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"lbm_(alpha|beta)_kernel|halo" -s 20 -c 6 --csv --log-file $O/ncu_small.csv \
   python tools/probe_overlap.py --axes x --only overlap --steps 8 > $O/ncu_small.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 10 > $O/bench_n1.json 2> $O/bench_n1.err
tail -4 $O/pytest_gpu.log; cut -c1-300 $O/p.jsonl; head -14 $O/timeline_xyz.txt; grep -v "^==" $O/ncu_small.csv | cut -d, -f5,13,15 | head -14; cut -c1-200 $O/bench_n1.json
