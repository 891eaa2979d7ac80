#!/bin/bash
# round 2c: BSK slice refilled in two halves (PBS5_SPLIT_BSK), lock step (kernel 5) and halves of the CTA half a step apart (51)
O=gpurun_out/lab35; mkdir -p $O
tools/lab/pbs_lab_s0 5 4 4096 2 > /dev/null 2>&1   # warm the clocks
for v in s0 s1; do for k in 5 51; do
  echo "{\"variant\": \"$v\", \"kernel\": $k}" >> $O/timing.jsonl
  timeout 120 tools/lab/pbs_lab_$v $k 4 592 3 | tail -1 >> $O/timing.jsonl 2>&1
  timeout 120 tools/lab/pbs_lab_$v $k 4 4096 3 | tail -1 >> $O/timing.jsonl 2>&1
done; done
cut -c1-150 $O/timing.jsonl
