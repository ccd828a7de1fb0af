#!/bin/bash
# Round-end verification on ONE B200 (run through gpurun): GPU parity tests, smoke, the default bench line, the
# reference arm, a sustained (>= 6 s) line with the clock record, a line with the STORE instantiations, the ncu launch
# list of the same bench command and one `--set full` capture of the two step kernels.  Everything lands in gpurun_out/check/.
O=gpurun_out/check; mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 400 python bench.py --min-seconds 6 --no-cpu-baseline > $O/bench_n1_sustained.json 2> $O/bench_n1_sustained.err
timeout 400 python bench.py --store --no-cpu-baseline > $O/bench_n1_store.json 2> $O/bench_n1_store.err
timeout 400 python bench.py --config 2 --no-cpu-baseline --steps 100 > $O/bench_n1_512.json 2> $O/bench_n1_512.err
timeout 400 python bench.py --config 4 --no-cpu-baseline --steps 100 > $O/bench_n1_f64_384.json 2> $O/bench_n1_f64_384.err
timeout 400 python bench.py --cs 0 --no-cpu-baseline > $O/bench_n1_bgk.json 2> $O/bench_n1_bgk.err
LBM_B200_PROFILE=1 timeout 300 python bench.py --no-cpu-baseline --no-verify > $O/bench_n1_profiled.json 2> $O/bench_n1_profiled.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
	python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > $O/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lbm_(alpha|beta)_kernel" -s 6 -c 2 -o $O/ncu_full_step_kernels \
	python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > $O/ncu_full.log 2>&1
ncu -i $O/ncu_full_step_kernels.ncu-rep --page raw --csv > $O/ncu_full_step_kernels_raw.csv 2>/dev/null
tail -3 $O/pytest_gpu.log; cat $O/smoke.log | tail -2; python tools/summarise_bench.py $O/bench_n1*.json; cut -c1-300 $O/bench_ref.json
