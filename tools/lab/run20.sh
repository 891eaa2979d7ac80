#!/bin/bash
O=gpurun_out/lab23; mkdir -p $O
for v in r0w0 r1w0 r0w1 r1w1; do
  echo "{\"variant\": \"$v\"}" >> $O/timing.jsonl
  tools/lab/pbs_lab_$v 5 4 592 3 | tail -1 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab_$v 5 4 4096 3 | tail -1 >> $O/timing.jsonl 2>&1
  tools/lab/pbs_lab_$v 5 3 444 3 | tail -1 >> $O/timing.jsonl 2>&1
done
cat $O/timing.jsonl
python -m pytest tests/test_generic_params.py tests/test_boolean.py tests/test_param_sets.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -15
