#!/bin/bash
# development: kernel5 vs kernel3 timing, bit difference, phase timeline
O=gpurun_out/lab8; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
for k in 3 5; do
  tools/lab/pbs_lab $k 4 4096 3 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab $k 4 592 3 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab $k 3 444 3 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab $k 2 296 3 >> $O/timing.jsonl 2>&1
done
tools/lab/pbs_lab_tl 5 4 592 1 $O/tl5_cts4.txt >> $O/tl.log 2>&1
tools/lab/pbs_lab_tl 5 2 296 1 $O/tl5_cts2.txt >> $O/tl.log 2>&1
cat $O/timing.jsonl
