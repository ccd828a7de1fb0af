#!/bin/bash
# round 2, job k: full GPU suite on the shipped build, smoke, proxies + timelines, ncu of the XFUSE step kernels, 1-GPU bench
O=gpurun_out/r2k; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
for a in "--axes x" "--axes xyz" "--axes xyz --size 512 --steps 40" "--axes x --size 512 --steps 40" "--axes x --size 1024x1024x32 --steps 100" "--axes x --size 384 --dtype f64 --steps 60" "--axes z" "--axes yz"; do
  timeout 120 python tools/probe_overlap.py $a >> $O/p.jsonl 2>> $O/p.err
done
timeout 120 python tools/probe_timeline.py x > $O/timeline_x.txt 2>&1
timeout 120 python tools/probe_timeline.py xyz > $O/timeline_xyz.txt 2>&1
timeout 120 python tools/probe_timeline.py xyz 512x512x512 > $O/timeline_xyz512.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lbm_(alpha|beta)_kernel" -s 20 -c 2 -o $O/xfuse_x256 \
   python tools/probe_overlap.py --axes x --only overlap --steps 8 > $O/ncu_xfuse.log 2>&1
ncu -i $O/xfuse_x256.ncu-rep --page raw --csv > $O/xfuse_x256.raw.csv 2>/dev/null
timeout 300 python bench.py --steps 200 --warmup 10 > $O/bench_n1.json 2> $O/bench_n1.err
tail -4 $O/pytest_gpu.log; cat $O/smoke.log | tail -2; cut -c1-60,150-330 $O/p.jsonl; cut -c1-200 $O/bench_n1.json
