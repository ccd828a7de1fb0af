#!/bin/bash
# round 2, job j: A/B of the XFUSE lane code -- default build, 3-D grid experiment, previous commit
O=gpurun_out/r2j; mkdir -p $O
L=turbulent_lbm_multigpu_b200/lib
for v in default grid3d old; do
  lib=$L/liblbm_b200.so; [ $v != default ] && lib=$L/liblbm_b200_$v.so
  for m in full nopush nopull none; do
    e=""; [ $m != full ] && e=$m
    echo "{\"variant\": \"$v\", \"mode\": \"$m\"}" >> $O/p.jsonl
    LBM_B200_LIB=$lib LBM_B200_XFUSE_DEBUG=$e timeout 120 python tools/probe_overlap.py --axes x >> $O/p.jsonl 2>> $O/p.err
  done
  echo "{\"variant\": \"$v\", \"mode\": \"xyz256 / x512 / x1024x1024x32\"}" >> $O/p.jsonl
  LBM_B200_LIB=$lib timeout 120 python tools/probe_overlap.py --axes xyz >> $O/p.jsonl 2>> $O/p.err
  LBM_B200_LIB=$lib timeout 120 python tools/probe_overlap.py --axes x --size 512 --steps 40 >> $O/p.jsonl 2>> $O/p.err
  LBM_B200_LIB=$lib timeout 120 python tools/probe_overlap.py --axes x --size 1024x1024x32 --steps 100 >> $O/p.jsonl 2>> $O/p.err
done
timeout 600 python -m pytest tests/test_gpu_multidomain.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
cut -c1-60,150-330 $O/p.jsonl; tail -3 $O/pytest_gpu.log
