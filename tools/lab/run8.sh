#!/bin/bash
O=gpurun_out/lab9; mkdir -p $O
for b in 4096 1184 1180 2000; do tools/lab/pbs_lab 5 4 $b 2 >> $O/timing.jsonl 2>&1; done
tools/lab/pbs_lab 5 3 1000 2 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab 3 4 4096 2 >> $O/timing.jsonl 2>&1
cat $O/timing.jsonl
