#!/bin/bash
cp tfhe_rs_string_b200/libb200tfhe.so /tmp/keep.so
for d in 1 2 4; do
  if [ $d != 1 ]; then cp tfhe_rs_string_b200/libb200tfhe_div$d.so tfhe_rs_string_b200/libb200tfhe.so; fi
  echo "DIV=$d"; python -m pytest tests/test_generic_params.py tests/test_boolean.py -m gpu -x -q -s 2>&1 | grep "KS+PBS in\|passed\|failed"
done
cp /tmp/keep.so tfhe_rs_string_b200/libb200tfhe.so
