#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/enc_time.py --batch 16384 --steps 10
timeout 300 python tools/enc_time.py --batch 16384 --steps 10 --pacbio
timeout 120 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_|pool3x3|stem_patch|tail" -s 109 -c 109 --csv --log-file gpurun_out/launches_cnn.csv python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > /dev/null 2>&1
