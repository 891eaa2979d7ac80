#!/bin/bash
O=gpurun_out/lab22; mkdir -p $O
tools/lab/pbs_lab_tl 72 1 4 1 $O/tl_lat72_b4.txt > $O/tl.log 2>&1
tools/lab/pbs_lab_tl 72 1 148 1 $O/tl_lat72_b148.txt >> $O/tl.log 2>&1
tools/lab/pbs_lab_tl 7 2 296 1 $O/tl_lat7_b296.txt >> $O/tl.log 2>&1
cat $O/tl.log
