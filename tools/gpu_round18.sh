#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_gpu.py -x -q -m gpu > gpurun_out/pytest_18.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_18.log
echo "== default"; timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== patch stem"; DVB_CNN_STEM_FUSED=0 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== pools all tiled"; DVB_CNN_POOL_TILED=2 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== pools old"; DVB_CNN_POOL_TILED=0 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== chunk 4096"; timeout 200 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_|pool3x3|stem_|tail" -c 330 --csv --log-file gpurun_out/launches_cnn.csv python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu exit $?"
