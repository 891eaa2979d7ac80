#!/bin/bash
O=gpurun_out/lab39; mkdir -p $O
tools/lab/pbs_lab 5 4 4096 2 > /dev/null 2>&1
for cfg in "4 592" "4 4096" "3 444" "2 296" "1 148"; do timeout 120 tools/lab/pbs_lab 5 $cfg 3 | tail -1 >> $O/timing.jsonl 2>&1; done
cut -c1-150 $O/timing.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; cut -c1-200 gpurun_out/r2c_bench_n1.json
