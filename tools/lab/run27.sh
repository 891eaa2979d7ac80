#!/bin/bash
O=gpurun_out/lab30; mkdir -p $O
tools/lab/pbs_lab 5 4 4096 3 > /dev/null 2>&1   # warm the clocks
for b in 1 16 148; do tools/lab/pbs_lab 74 1 $b 5 | tail -1 >> $O/timing.jsonl 2>&1; done
cut -c1-72 $O/timing.jsonl
tools/lab/pbs_lab_tl 74 1 4 1 $O/tl_lat4_b4.txt > $O/tl.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pbs_lat4_kernel -c 1 -o $O/pbs_lat4 tools/lab/pbs_lab 74 1 16 1 > $O/ncu.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
