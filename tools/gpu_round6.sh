#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cnn_probe.py 3 > gpurun_out/probe.log 2>&1; echo "probe exit $?"; tail -2 gpurun_out/probe.log | cut -c1-100
for kb in 40 56 72 108; do for ms in 3 4 6; do echo "SMEM_KB=$kb MAX_STAGES=$ms"; DVB_CNN_SMEM_KB=$kb DVB_CNN_MAX_STAGES=$ms timeout 120 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3; done; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_gemm|pool3x3|stem_patch|tail" -s 109 -c 109 --csv --log-file gpurun_out/launches_cnn.csv python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
