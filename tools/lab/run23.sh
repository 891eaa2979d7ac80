#!/bin/bash
O=gpurun_out/lab26; mkdir -p $O
tools/lab/pbs_lab 5 4 4096 3 > /dev/null 2>&1   # warm the clocks
for b in 1 4 16 32 64 96 128 148; do
  tools/lab/pbs_lab 72 1 $b 5 | tail -1 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab 74 1 $b 5 | tail -1 >> $O/timing.jsonl 2>&1
done
cut -c1-70 $O/timing.jsonl
