#!/bin/bash
# round 2, GPU job A (1 GPU): new parity tests on the round-1 kernels + overlap-cost proxy baseline
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multiprocess.py::test_two_processes_sharing_one_gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -m pytest tests/test_gpu_multiprocess.py -m gpu -q -k two_processes > $O/pytest_shared.log 2>&1; echo "rc=$?" >> $O/pytest_shared.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
for ax in z y x yz xyz; do
  timeout 120 python tools/probe_overlap.py --axes $ax --timeline >> $O/proxy_256.jsonl 2>> $O/proxy.err
done
timeout 120 python tools/probe_overlap.py --axes z --size 512x512x64 --timeline >> $O/proxy_strong.jsonl 2>> $O/proxy.err
timeout 120 python tools/probe_overlap.py --axes xyz --size 512 --steps 40 --timeline >> $O/proxy_512.jsonl 2>> $O/proxy.err
timeout 120 python tools/probe_overlap.py --axes z --size 384x384x48 --dtype f64 --timeline >> $O/proxy_strong.jsonl 2>> $O/proxy.err
tail -3 $O/pytest_gpu.log; tail -3 $O/pytest_shared.log; tail -2 $O/smoke.log; cat $O/proxy_*.jsonl
