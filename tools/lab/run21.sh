#!/bin/bash
O=gpurun_out/lab24; mkdir -p $O
for v in l0 l1; do
  echo "{\"variant\": \"$v\"}" >> $O/timing.jsonl
  tools/lab/pbs_lab_$v 72 1 4 5 | tail -1 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab_$v 72 1 148 5 | tail -1 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab_$v 7 2 296 5 | tail -1 >> $O/timing.jsonl 2>&1
done
cat $O/timing.jsonl
