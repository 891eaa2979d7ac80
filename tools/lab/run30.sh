#!/bin/bash
O=gpurun_out/lab33; mkdir -p $O
tools/lab/pbs_lab_b0 5 4 4096 3 > /dev/null 2>&1   # warm the clocks
for b in 150 296; do for v in b0 b1; do tools/lab/pbs_lab_$v 7 2 $b 5 | tail -1 >> $O/timing.jsonl 2>&1; done; done
for v in b0 b1; do tools/lab/pbs_lab_$v 72 1 148 5 | tail -1 >> $O/timing.jsonl 2>&1; done
cut -c1-72 $O/timing.jsonl
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
