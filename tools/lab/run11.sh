#!/bin/bash
O=gpurun_out/lab12; mkdir -p $O
for c in 1 2 3 4; do tools/lab/pbs_lab 5 $c $((148*c)) 3 | tail -1 >> $O/timing.jsonl; tools/lab/pbs_lab 3 $c $((148*c)) 3 | tail -1 >> $O/timing.jsonl; done
tools/lab/pbs_lab_tl 5 4 592 1 $O/tl5_cts4.txt >> $O/tl.log 2>&1
tools/lab/pbs_lab_tl 5 3 444 1 $O/tl5_cts3.txt >> $O/tl.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pbs_kernel5 -c 1 -o $O/k5_cts4 tools/lab/pbs_lab 5 4 592 1 > $O/ncu_k5.log 2>&1
cat $O/timing.jsonl
