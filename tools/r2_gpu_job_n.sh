#!/bin/bash
# round 2, job n: fused x exchange on rows that are / are not multiples of 128 bytes
O=gpurun_out/r2n; mkdir -p $O
for a in "--axes x --size 300x256x256" "--axes x --size 320x256x256" "--axes x --size 288x256x256" "--axes x --size 296x256x256"; do
  timeout 120 python tools/probe_overlap.py $a >> $O/p.jsonl 2>> $O/p.err
  LBM_B200_XFUSE=0 timeout 120 python tools/probe_overlap.py $a >> $O/p_nofuse.jsonl 2>> $O/p.err
done
cut -c1-60,100-330 $O/p.jsonl; echo; cut -c1-60,100-330 $O/p_nofuse.jsonl
