#!/bin/bash
# round 2c: from_torus in three FP64 instructions (f) and the forward twist folded into the first pass (t)
O=gpurun_out/lab20; mkdir -p $O
for v in f0t0 f1t0 f0t1 f1t1; do
  echo "{\"variant\": \"$v\"}" >> $O/timing.jsonl
  tools/lab/pbs_lab_$v 5 4 592 3 | tail -2 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab_$v 5 4 4096 3 | tail -1 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab_$v 5 3 444 3 | tail -1 >> $O/timing.jsonl 2>&1
done
cat $O/timing.jsonl
