#!/bin/bash
O=gpurun_out/lab36; mkdir -p $O
tools/lab/pbs_lab_s0 5 4 4096 2 > /dev/null 2>&1
tools/lab/pbs_lab_s0_tl 5 4 592 1 $O/tl_s0.txt > $O/tl.log 2>&1
tools/lab/pbs_lab_s1_tl 5 4 592 1 $O/tl_s1.txt >> $O/tl.log 2>&1
tools/lab/pbs_lab_s1 5 3 444 3 | tail -1
tools/lab/pbs_lab_s0 5 3 444 3 | tail -1
