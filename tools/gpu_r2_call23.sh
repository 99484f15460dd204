#!/bin/bash
# Round 2, GPU call 23: the failing plan-2 test on the tree of call 18 (commit 9aa1ef2, built into _old/) and on HEAD, alone, twice each.
mkdir -p gpurun_out
for i in 1 2; do (cd _old && timeout 300 python -m pytest tests/test_cnn_gpu.py -m gpu -q -p no:cacheprovider -k wider > ../gpurun_out/c23_old_$i.log 2>&1; echo "old $i exit $?: $(grep -E 'AssertionError|passed|failed' ../gpurun_out/c23_old_$i.log | tr '\n' ' ')"); done
for i in 1 2; do timeout 300 python -m pytest tests/test_cnn_gpu.py -m gpu -q -p no:cacheprovider -k wider > gpurun_out/c23_new_$i.log 2>&1; echo "new $i exit $?: $(grep -E 'AssertionError|passed|failed' gpurun_out/c23_new_$i.log | tr '\n' ' ')"; done
(cd _old && timeout 600 python -m pytest tests/test_cnn_gpu.py -m gpu -q -p no:cacheprovider > ../gpurun_out/c23_old_file.log 2>&1; echo "old file exit $?: $(grep -E 'AssertionError|passed|failed' ../gpurun_out/c23_old_file.log | tr '\n' ' ')")
