#!/bin/bash
O=gpurun_out/lab7; mkdir -p $O
tools/lab/pbs_lab_f0 3 4 4096 3 > $O/timing.jsonl 2>&1
tools/lab/pbs_lab_f1 3 4 4096 3 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab_f1 3 3 444 3 >> $O/timing.jsonl 2>&1
tools/lab/pbs_lab_tl 3 4 592 1 $O/tl_fused.txt >> $O/tl.log 2>&1
cat $O/timing.jsonl
