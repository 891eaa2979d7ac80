#!/bin/bash
O=gpurun_out/lab11; mkdir -p $O
for f in 0 1 2 3; do
  tools/lab/pbs_lab_f$f 5 4 4096 3 | tail -1 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab_f$f 5 4 592 3 | tail -1 >> $O/timing.jsonl 2>&1
done
tools/lab/pbs_lab_f3 5 3 444 3 | tail -1 >> $O/timing.jsonl
tools/lab/pbs_lab_f3 5 2 296 3 | tail -1 >> $O/timing.jsonl
cat $O/timing.jsonl
