#!/bin/bash
# A/B of the two latency kernels inside the program executor: same box, same process order
O=gpurun_out/lab27; mkdir -p $O
python tools/perf_single_ops.py > $O/warm.jsonl 2>&1
python tools/perf_single_ops.py > $O/lat4.jsonl 2>&1
cp tfhe_rs_string_b200/libb200tfhe.so /tmp/keep.so; cp tfhe_rs_string_b200/libb200tfhe_lat2.so tfhe_rs_string_b200/libb200tfhe.so
python tools/perf_single_ops.py > $O/lat2.jsonl 2>&1
cp /tmp/keep.so tfhe_rs_string_b200/libb200tfhe.so
python tools/perf_single_ops.py > $O/lat4_again.jsonl 2>&1
paste <(cut -c1-200 $O/lat4.jsonl | python -c "import sys,json; [print(json.loads(l)['op'], json.loads(l)['ms']) for l in sys.stdin]") <(python -c "import sys,json; [print(json.loads(l)['ms']) for l in open('$O/lat2.jsonl')]") <(python -c "import sys,json; [print(json.loads(l)['ms']) for l in open('$O/lat4_again.jsonl')]")
