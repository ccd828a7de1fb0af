#!/bin/bash
# round 2, job l: fused x exchange on any row length (partly empty last block, block size that divides a row)
O=gpurun_out/r2l; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
for a in "--axes x" "--axes x --size 384 --dtype f64 --steps 60" "--axes xyz --size 384 --dtype f64 --steps 60" "--axes x --size 300x256x256" "--axes x --size 384 --steps 60" "--axes xyz"; do
  timeout 120 python tools/probe_overlap.py $a >> $O/p.jsonl 2>> $O/p.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lbm_(alpha|beta)_kernel" -s 12 -c 2 -o $O/xfuse_x256 \
   python tools/probe_overlap.py --axes x --only overlap --steps 8 > $O/ncu_xfuse.log 2>&1
ncu -i $O/xfuse_x256.ncu-rep --page raw --csv > $O/xfuse_x256.raw.csv 2>/dev/null
rm -f $O/xfuse_x256.ncu-rep
tail -4 $O/pytest_gpu.log; tail -1 $O/smoke.log; cut -c1-60,150-330 $O/p.jsonl; tail -3 $O/ncu_xfuse.log
