#!/bin/bash
O=gpurun_out/lab19; mkdir -p $O
for d in 0 1000 2000 3000 4000 4500 5000 6000; do
  echo "{\"delay_per_ct\": $d}" >> $O/timing.jsonl
  LAB_DELAY=$d tools/lab/pbs_lab_delay 5 4 592 3 | tail -1 >> $O/timing.jsonl 2>&1
  LAB_DELAY=$d tools/lab/pbs_lab_delay 5 4 4096 2 | tail -1 >> $O/timing.jsonl 2>&1
done
for d in 0 4000; do LAB_DELAY=$d tools/lab/pbs_lab_delay_tl 5 4 592 1 $O/tl_d$d.txt >> $O/tl.log 2>&1; done
cat $O/timing.jsonl
