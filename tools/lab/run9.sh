#!/bin/bash
O=gpurun_out/lab10; mkdir -p $O
tools/mb/tmem_bw > $O/tmem_bw.jsonl 2>&1
tools/lab/pbs_lab_nosat 5 4 4096 2 > $O/nosat.jsonl 2>&1
tools/lab/pbs_lab_nosat 5 3 1000 2 >> $O/nosat.jsonl 2>&1
tools/lab/pbs_lab_nosat 5 1 300 2 >> $O/nosat.jsonl 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pbs_kernel5 -c 1 -o $O/k5_cts4 tools/lab/pbs_lab 5 4 592 1 > $O/ncu_k5.log 2>&1
cat $O/tmem_bw.jsonl $O/nosat.jsonl; tail -3 $O/ncu_k5.log
