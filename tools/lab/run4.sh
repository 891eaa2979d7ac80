#!/bin/bash
O=gpurun_out/lab5; mkdir -p $O
tools/lab/pbs_lab 3 4 4096 3 > $O/timing.jsonl 2>&1
tools/lab/pbs_lab_dup 3 4 4096 3 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab_dup 31 4 4096 3 >> $O/timing.jsonl 2>&1
cat $O/timing.jsonl
