#!/bin/bash
# development: ncu launch list + full capture (source-level) of the dominant kernel and of the latency kernel, after the plain run exited 0
O=gpurun_out/prof_r02c; mkdir -p $O
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err || exit 1
python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain.json 2> $O/plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:pbs_kernel5 -c 1 -o $O/pbs_kernel5 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:pbs_lat4_kernel -c 1 -o $O/pbs_lat4 python -c "import __graft_entry__ as g; g.smoke()" > $O/ncu_lat4.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:ks_mma_kernel -c 1 -o $O/ks_mma python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_ks.log 2>&1
ls -la $O
