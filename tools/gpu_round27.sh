#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_gpu.py -x -q -m gpu > gpurun_out/pytest_27.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_27.log
echo "== s4 padded to 96"; timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
echo "== no padding"; DVB_CNN_PAD_CIN32=0 timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
echo "== pad also 48-channel tensors (5x5 inputs) to 64"; DVB_CNN_PAD_CIN32_MIN=48 timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
echo "== precise"; timeout 200 python tools/cnn_time.py --batch 4096 --chunk 2048 --steps 3 --precision 1
