#!/bin/bash
O=gpurun_out/lab13; mkdir -p $O
for o in 0 1; do
  tools/lab/pbs_lab_o$o 5 4 4096 3 | tail -1 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab_o$o 5 4 592 3 | tail -1 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab_o$o 5 3 444 3 | tail -1 >> $O/timing.jsonl 2>&1
done
cat $O/timing.jsonl
