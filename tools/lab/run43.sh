#!/bin/bash
O=gpurun_out/lab47; mkdir -p $O
tools/lab/pbs_lab_p1 5 4 4096 2 > /dev/null 2>&1
for rep in 1 2 3; do for v in p1 p2; do
  timeout 120 tools/lab/pbs_lab_$v 5 4 4096 3 | tail -1 | cut -c1-75 | sed "s/^/$v /" >> $O/timing.txt
  timeout 120 tools/lab/pbs_lab_$v 5 3 444 3 | tail -1 | cut -c1-75 | sed "s/^/$v /" >> $O/timing.txt
done; done
cat $O/timing.txt
