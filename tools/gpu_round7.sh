#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cnn_probe.py 3 > gpurun_out/probe.log 2>&1; echo "probe exit $?"; grep -E "rel_err" gpurun_out/probe.log | awk '{print $6, $1}' | sort -g -r | head -3; tail -2 gpurun_out/probe.log | cut -c1-100
timeout 120 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_gemm|pool3x3|stem_patch|tail" -s 109 -c 109 --csv --log-file gpurun_out/launches_cnn.csv python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
