#!/bin/bash
# round 2, GPU job C (1 GPU): fused x exchange (push + lazy pull inside the step kernels)
O=gpurun_out/r2c; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
for ax in z x xyz; do
  timeout 120 python tools/probe_overlap.py --axes $ax --timeline >> $O/proxy_256.jsonl 2>> $O/proxy.err
done
timeout 120 python tools/probe_overlap.py --axes xyz --size 512 --steps 40 --timeline >> $O/proxy_512.jsonl 2>> $O/proxy.err
timeout 120 python tools/probe_overlap.py --axes x --size 1024x1024x32 --steps 100 --timeline >> $O/proxy_recipe.jsonl 2>> $O/proxy.err
timeout 120 python tools/probe_overlap.py --axes x --size 384 --dtype f64 --steps 60 --timeline >> $O/proxy_f64.jsonl 2>> $O/proxy.err
tail -5 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/proxy_*.jsonl; tail -5 $O/proxy.err
