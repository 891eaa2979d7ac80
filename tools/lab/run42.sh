#!/bin/bash
O=gpurun_out/lab46; mkdir -p $O
tools/lab/pbs_lab_ls0 5 4 4096 2 > /dev/null 2>&1
for v in ls0 ls1; do
  echo "{\"variant\": \"$v\"}" >> $O/timing.jsonl
  for cfg in "7 2 150" "7 2 296"; do timeout 120 tools/lab/pbs_lab_$v $cfg 5 | tail -1 >> $O/timing.jsonl 2>&1; done
done
cut -c1-110 $O/timing.jsonl
