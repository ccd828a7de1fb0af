#!/bin/bash
# round 2, final check of the committed build on one GPU: what the driver runs at round end
O=gpurun_out/r2final; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_n1_driver_like.json 2> $O/bench_n1_driver_like.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 300 python bench.py --steps 400 --warmup 20 > $O/bench_n1.json 2> $O/bench_n1.err
tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log; cut -c1-260 $O/bench_n1_driver_like.json; cut -c1-200 $O/bench_ref.json; cut -c1-200 $O/bench_n1.json
