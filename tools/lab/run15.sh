#!/bin/bash
O=gpurun_out/lab17; mkdir -p $O
tools/lab/pbs_lab 7 1 148 3 | tail -1 >> $O/timing.jsonl
tools/lab/pbs_lab 7 2 296 3 | tail -1 >> $O/timing.jsonl
tools/lab/pbs_lab 7 1 4 3 | tail -1 >> $O/timing.jsonl
tools/lab/pbs_lab_tl 7 1 148 1 $O/tl_lat1.txt >> $O/tl.log 2>&1
tools/lab/pbs_lab_tl 7 2 296 1 $O/tl_lat2.txt >> $O/tl.log 2>&1
cat $O/timing.jsonl
