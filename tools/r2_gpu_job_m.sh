#!/bin/bash
# round 2, job m: shared-GPU rank threads with 32 hardware queues, block size of the fused launches by idle threads
O=gpurun_out/r2m; mkdir -p $O
timeout 600 python -m pytest tests/test_host_cpp.py -m gpu -q -x > $O/pytest_host_cpp.log 2>&1; echo "rc=$?" >> $O/pytest_host_cpp.log
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 300 python -m pytest tests/test_host_cpp.py -m gpu -q -x -k "one_sided" > $O/pytest_host_cpp_8queues.log 2>&1; echo "rc=$?" >> $O/pytest_host_cpp_8queues.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_host_cpp.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
for a in "--axes x --size 300x256x256" "--axes x --size 384 --dtype f64 --steps 60" "--axes x"; do
  timeout 120 python tools/probe_overlap.py $a >> $O/p.jsonl 2>> $O/p.err
done
tail -3 $O/pytest_host_cpp.log; tail -3 $O/pytest_host_cpp_8queues.log; tail -3 $O/pytest_gpu.log; cut -c1-60,150-330 $O/p.jsonl
