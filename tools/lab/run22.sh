#!/bin/bash
O=gpurun_out/lab25; mkdir -p $O
for k in "72 1 4" "74 1 4" "72 1 148" "74 1 148" "74 1 1" "74 1 64"; do
  tools/lab/pbs_lab $k 5 | tail -1 >> $O/timing.jsonl 2>&1
done
tools/lab/pbs_lab_tl 74 1 4 1 $O/tl_lat4_b4.txt > $O/tl.log 2>&1
tools/lab/pbs_lab_tl 74 1 148 1 $O/tl_lat4_b148.txt >> $O/tl.log 2>&1
cat $O/timing.jsonl
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
