#!/bin/bash
O=gpurun_out/lab6; mkdir -p $O
tools/lab/pbs_lab 3 2 296 3 > $O/timing.jsonl 2>&1
tools/lab/pbs_lab_dup2 3 2 296 3 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab 3 4 592 3 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab_dup2 3 4 592 3 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab_dup 3 4 592 3 >> $O/timing.jsonl 2>&1
cat $O/timing.jsonl
