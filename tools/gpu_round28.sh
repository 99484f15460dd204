#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_gpu.py -x -q -m gpu > gpurun_out/pytest_28.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_28.log
echo "== default (48->64, 80->96)"; timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
for pct in 12 20 34 60; do echo "== pad64 maxpct $pct"; DVB_CNN_PAD_CIN64_MAXPCT=$pct timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3; done
DVB_CNN_PAD_CIN64_MAXPCT=34 timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q -m gpu -k "block_outputs or branch or pacbio" 2>&1 | tail -2
