#!/bin/bash
O=gpurun_out/lab43; mkdir -p $O
tools/lab/pbs_lab_k1 5 4 4096 2 > /dev/null 2>&1
for v in k1 k3; do
  echo "{\"variant\": \"$v\"}" >> $O/timing.jsonl
  for cfg in "4 592" "4 4096" "3 444" "4 700"; do timeout 120 tools/lab/pbs_lab_$v 5 $cfg 3 | tail -1 >> $O/timing.jsonl 2>&1; done
done
cut -c1-150 $O/timing.jsonl
