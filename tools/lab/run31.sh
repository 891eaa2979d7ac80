#!/bin/bash
O=gpurun_out/lab34; mkdir -p $O
tools/lab/pbs_lab 5 4 4096 3 > /dev/null 2>&1   # warm the clocks
for b in 1 16 148; do for v in pbs_lab pbs_lab_p1; do tools/lab/$v 74 1 $b 5 | tail -1 >> $O/timing.jsonl 2>&1; done; done
cut -c1-72 $O/timing.jsonl
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
