#!/bin/bash
O=gpurun_out/lab4; mkdir -p $O
tools/lab/pbs_lab 3 4 4096 3 > $O/timing.jsonl 2>&1
tools/lab/pbs_lab 3 4 592 3 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab 3 2 296 3 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab_tl 3 4 592 1 $O/tl_cts4.txt >> $O/tl.log 2>&1
cat $O/timing.jsonl
