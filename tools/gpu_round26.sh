#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_golden.py tests/test_encoder_gpu.py tests/test_call_variants.py -x -q -m gpu > gpurun_out/pytest_26.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_26.log
echo "== merged 1x1"; timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
echo "== not merged"; DVB_CNN_MERGE_1X1=0 timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
echo "== merged, persist never"; DVB_CNN_PERSIST=0 timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_|pool|stem_|tail" -c 200 --csv --log-file gpurun_out/launches_cnn_merged.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu exit $?"
