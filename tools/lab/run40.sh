#!/bin/bash
# ptxas option variants: x0 default, x1 --allow-expensive-optimizations, x2 / x3 --register-usage-level 10 / 0
O=gpurun_out/lab44; mkdir -p $O
tools/lab/pbs_lab_x0 5 4 4096 2 > /dev/null 2>&1
for v in x0 x1 x2 x3; do
  echo "{\"variant\": \"$v\"}" >> $O/timing.jsonl
  for cfg in "5 4 4096" "5 3 444" "74 1 148"; do timeout 120 tools/lab/pbs_lab_$v $cfg 3 | tail -1 >> $O/timing.jsonl 2>&1; done
done
cut -c1-110 $O/timing.jsonl
