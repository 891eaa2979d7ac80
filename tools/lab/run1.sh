#!/bin/bash
# development: GPU experiment batch 1 (single-warp phase timings, pipe model)
set -x
O=gpurun_out/lab1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
tools/mb/pipes > $O/pipes.jsonl 2>&1
for c in 1 2 3 4; do tools/lab/pbs_lab 3 $c $((148*c)) 3 >> $O/lab_timing.jsonl 2>&1; done
tools/lab/pbs_lab 3 4 4096 3 >> $O/lab_timing.jsonl 2>&1
for c in 1 2 4; do tools/lab/pbs_lab_tl 3 $c $((148*c)) 1 $O/tl_cts$c.txt >> $O/lab_tl.log 2>&1; done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pbs_kernel3 -c 1 -o $O/k3_cts2 tools/lab/pbs_lab 3 2 296 1 > $O/ncu_cts2.log 2>&1
ls -la $O
