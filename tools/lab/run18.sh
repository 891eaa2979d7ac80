#!/bin/bash
O=gpurun_out/lab21; mkdir -p $O
tools/lab/pbs_lab_tl 5 4 592 1 $O/tl_k5_cts4.txt > $O/tl.log 2>&1
tools/lab/pbs_lab_tl 5 1 148 1 $O/tl_k5_cts1.txt >> $O/tl.log 2>&1
tools/lab/pbs_lab_tl 5 2 296 1 $O/tl_k5_cts2.txt >> $O/tl.log 2>&1
N="A,fwdFFT,waits,-,ownmul,barfull,othmul+inv1,bskissue,invrest,D"
for c in 4 1 2; do python tools/timeline.py $O/tl_k5_cts$c.txt $N | head -20; done
