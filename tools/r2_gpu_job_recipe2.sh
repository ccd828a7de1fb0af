#!/bin/bash
# round 2, 4 GPUs, after the lane-code rework: the reference's own benchmark recipes (see r2_gpu_job_recipe.sh), Smagorinsky
R=$PWD
for mode in weak-1d strong-1d; do
  for cs in 0.1; do
    D=$R/gpurun_out/r2recipe2/${mode}_cs${cs}; mkdir -p $D; cd $D
    MAX=4; [ $mode = strong-1d ] && MAX=2
    timeout 600 python $R/tools/benchmark.py $mode --max-num $MAX --num-exp 2 --reference-recipe --loops 300 --smagorinsky $cs > run.log 2>&1
    echo "== $mode cs=$cs"; tail -6 run.log
    cd $R
  done
done
