#!/bin/bash
# Round 2, GPU call 22: how often, and under which plan, does test_wider_cta_pair_rule_equals_default_plan fail (seen once in call 20)?
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_cnn_gpu.py -m gpu -q -p no:cacheprovider -k wider > gpurun_out/c22_alone_$i.log 2>&1; echo "alone $i exit $?: $(grep -E 'AssertionError|passed|failed' gpurun_out/c22_alone_$i.log | tr '\n' ' ')"; done
for i in 1 2 3; do timeout 600 python -m pytest tests/test_cnn_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/c22_file_$i.log 2>&1; echo "file $i exit $?: $(grep -E 'AssertionError|passed|failed' gpurun_out/c22_file_$i.log | tr '\n' ' ')"; done
