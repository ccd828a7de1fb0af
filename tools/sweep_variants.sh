for v in "" _lb256x2 _lb128x5 _lb128x6 _lb128x8; do
  echo "=== variant ${v:-default}"
  LBM_B200_LIB=$PWD/turbulent_lbm_multigpu_b200/lib/liblbm_b200$v.so python tools/probe_perf.py --size 256 --steps 30 --cs 0.0 0.1 --block 64 128 --modes alpha beta
done
