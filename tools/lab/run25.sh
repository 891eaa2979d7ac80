#!/bin/bash
O=gpurun_out/lab28; mkdir -p $O
tools/lab/pbs_lab 5 4 4096 3 > /dev/null 2>&1   # warm the clocks
for b in 1 16 148; do
  for k in 72 74 75; do tools/lab/pbs_lab $k 1 $b 5 | tail -1 >> $O/timing.jsonl 2>&1; done
done
cut -c1-72 $O/timing.jsonl
tools/lab/pbs_lab_tl 75 1 4 1 $O/tl_lat4t_b4.txt > $O/tl.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
