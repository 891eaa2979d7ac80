#!/bin/bash
O=gpurun_out/lab31; mkdir -p $O
tools/lab/pbs_lab_c0 5 4 4096 3 > /dev/null 2>&1   # warm the clocks
for b in 1 16 148; do for v in c0 c1; do tools/lab/pbs_lab_$v 74 1 $b 5 | tail -1 >> $O/timing.jsonl 2>&1; done; done
cut -c1-72 $O/timing.jsonl
