#!/bin/bash
# Round 2, GPU call 24: the pool-source padding fix (plan-2 test three times, then the classifier test file), then the driver-shaped bench
# line of the ABI-v4 code and the reference arm.
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python -m pytest tests/test_cnn_gpu.py -m gpu -q -p no:cacheprovider -k wider > gpurun_out/c24_wider_$i.log 2>&1; echo "wider $i exit $?: $(grep -E 'AssertionError|passed|failed' gpurun_out/c24_wider_$i.log | tr '\n' ' ')"; done
timeout 900 python -m pytest tests/test_cnn_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/c24_cnn_file.log 2>&1; echo "cnn file exit $?: $(tail -1 gpurun_out/c24_cnn_file.log)"
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/c24_bench_n1.json 2> gpurun_out/c24_bench_err.txt; echo "bench exit $?"; tail -c 600 gpurun_out/c24_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c24_bench_reference.json 2> gpurun_out/c24_bench_ref_err.txt; echo "reference arm exit $?"; tail -c 400 gpurun_out/c24_bench_reference.json
