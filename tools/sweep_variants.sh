#!/bin/bash
# usage: tools/sweep_variants.sh [tag ...]   -- probe_perf.py over the default library and the named variants
for v in "" "$@"; do
  lib=$PWD/turbulent_lbm_multigpu_b200/lib/liblbm_b200${v:+_$v}.so
  echo "=== variant ${v:-default}"
  LBM_B200_LIB=$lib python tools/probe_perf.py --size ${SIZE:-256} --dtype ${DTYPE:-f32} --steps 40 --cs 0.1 --vw 2 --block 128 --modes alpha beta both
done
