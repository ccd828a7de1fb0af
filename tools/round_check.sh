#!/bin/bash
# Round-end verification on one B200 (run through gpurun): GPU parity tests, smoke, the default bench line,
# the reference arm, a profiled run, and the ncu launch list of the same bench command.
# Everything lands in gpurun_out/check/.
O=gpurun_out/check; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
LBM_B200_PROFILE=1 timeout 300 python bench.py --no-cpu-baseline > $O/bench_n1_profiled.json 2> $O/bench_n1_profiled.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
	python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
tail -3 $O/pytest_gpu.log; cat $O/smoke.log | tail -2; cat $O/bench_n1.json | cut -c1-400
