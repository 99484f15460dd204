#!/bin/bash
mkdir -p gpurun_out
export DVB_CNN_PERSIST=0
echo "== non-persistent first: tests + timing (pool-after-conv, fused stem grid x4)"
timeout 900 python -m pytest tests/test_cnn_gpu.py -x -q -m gpu > gpurun_out/pytest_19a.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_19a.log
echo "== default (fused stem)"; timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== patch stem"; DVB_CNN_STEM_FUSED=0 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== patch stem, pool before conv"; DVB_CNN_STEM_FUSED=0 DVB_CNN_POOL_AFTER_CONV=0 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== patch stem, old pools"; DVB_CNN_STEM_FUSED=0 DVB_CNN_POOL_TILED=0 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
export DVB_CNN_PERSIST=1
echo "== PERSISTENT: tests"
timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q -m gpu -k "stem_layers or block_outputs or branch or pacbio or chunking" > gpurun_out/pytest_19b.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_19b.log
echo "== persistent, patch stem"; DVB_CNN_STEM_FUSED=0 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== persistent, patch stem, lanes 1"; DVB_CNN_LANES=1 DVB_CNN_STEM_FUSED=0 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== persistent, fused stem"; timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
DVB_CNN_STEM_FUSED=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_|pool3x3|stem_|tail" -c 330 --csv --log-file gpurun_out/launches_cnn_persist.csv python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu exit $?"
