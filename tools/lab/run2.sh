#!/bin/bash
O=gpurun_out/lab3; mkdir -p $O
tools/lab/pbs_lab 3 4 4096 3 > $O/timing.jsonl 2>&1
tools/lab/pbs_lab 31 4 4096 3 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab 31 4 592 3 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab_tl 31 4 592 1 $O/tl_dephase.txt >> $O/tl.log 2>&1
cat $O/timing.jsonl
