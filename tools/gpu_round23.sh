#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q -m gpu -k "stem or chunking or block_outputs" > gpurun_out/pytest_23.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_23.log
timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stem_conv1 -s 1 -c 1 -o gpurun_out/stem_full -f python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > gpurun_out/stem_full.log 2>&1; echo "ncu exit $?"; ls -la gpurun_out/stem_full.ncu-rep
