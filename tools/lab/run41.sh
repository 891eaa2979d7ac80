#!/bin/bash
O=gpurun_out/lab45; mkdir -p $O
tools/lab/pbs_lab_xt1 5 4 4096 2 > /dev/null 2>&1
for v in xt1 xt0; do
  echo "{\"variant\": \"$v\"}" >> $O/timing.jsonl
  for cfg in "5 4 592" "5 4 4096"; do timeout 120 tools/lab/pbs_lab_$v $cfg 3 | tail -1 >> $O/timing.jsonl 2>&1; done
done
cut -c1-110 $O/timing.jsonl
