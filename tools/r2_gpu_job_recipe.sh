#!/bin/bash
# round 2, 4 GPUs: the reference's own benchmark recipes (/root/reference/benchmark.py:62-83 weak: x = 1024 n, y = 1024,
# z = 32, -X n; :107-129 strong: 1024 x 1024 x 32, np = 2^k) through tools/benchmark.py and the C++ driver lbm_b200
# (one host thread per GPU), BGK like the reference and with Smagorinsky.
R=$PWD
for mode in weak-1d strong-1d; do
  for cs in 0 0.1; do
    D=$R/gpurun_out/r2recipe/${mode}_cs${cs}; mkdir -p $D; cd $D
    MAX=4; [ $mode = strong-1d ] && MAX=2
    EXTRA=""; [ $cs != 0 ] && EXTRA="--smagorinsky $cs"
    timeout 600 python $R/tools/benchmark.py $mode --max-num $MAX --num-exp 2 --reference-recipe --loops 300 $EXTRA > run.log 2>&1
    echo "== $mode cs=$cs"; tail -6 run.log
    cd $R
  done
done
