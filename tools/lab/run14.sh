#!/bin/bash
O=gpurun_out/lab15; mkdir -p $O
for k in 5 51; do
  tools/lab/pbs_lab $k 4 4096 3 | tail -1 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab $k 4 592 3 | tail -1 >> $O/timing.jsonl 2>&1
done
tools/lab/pbs_lab_tl 51 4 592 1 $O/tl51.txt >> $O/tl.log 2>&1
cat $O/timing.jsonl
