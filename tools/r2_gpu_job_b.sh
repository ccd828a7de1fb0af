#!/bin/bash
# round 2, GPU job B (1 GPU): device-resident halo sequence numbers, wait+unpack split, fused x push
O=gpurun_out/r2b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
for ax in z x xyz; do
  timeout 120 python tools/probe_overlap.py --axes $ax --timeline >> $O/proxy_256.jsonl 2>> $O/proxy.err
done
LBM_B200_XFUSE=0 timeout 120 python tools/probe_overlap.py --axes x --timeline >> $O/proxy_256_nofuse.jsonl 2>> $O/proxy.err
timeout 120 python tools/probe_overlap.py --axes xyz --size 512 --steps 40 --timeline >> $O/proxy_512.jsonl 2>> $O/proxy.err
timeout 120 python tools/probe_overlap.py --axes x --size 1024x1024x32 --steps 100 --timeline >> $O/proxy_recipe.jsonl 2>> $O/proxy.err
timeout 120 python tools/probe_overlap.py --axes z --size 512x512x64 --timeline >> $O/proxy_strong.jsonl 2>> $O/proxy.err
timeout 120 python tools/probe_overlap.py --axes z --size 384x384x48 --dtype f64 --timeline >> $O/proxy_strong.jsonl 2>> $O/proxy.err
timeout 300 python bench.py --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
tail -5 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/proxy_*.jsonl; cut -c1-300 $O/bench_n1.json; tail -5 $O/proxy.err
