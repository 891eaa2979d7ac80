#!/bin/bash
O=gpurun_out/lab14; mkdir -p $O
for d in 0 300 600 1000 1500 2500 4000 6000 9000 12000; do
  echo "{\"delay\": $d}" >> $O/timing.jsonl
  LAB_DELAY=$d tools/lab/pbs_lab_delay 5 4 592 3 | tail -1 >> $O/timing.jsonl 2>&1
done
for d in 600 2500 6000; do LAB_DELAY=$d tools/lab/pbs_lab_delay_tl 5 4 592 1 $O/tl_d$d.txt >> $O/tl.log 2>&1; done
cat $O/timing.jsonl
