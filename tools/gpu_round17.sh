#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_encoder_gpu.py tests/test_golden.py -x -q -m gpu > gpurun_out/pytest_17.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_17.log
timeout 300 python tools/enc_time.py --batch 16384 --steps 10
echo "== default (lanes 4, tiled pools, any-N tiles)"; timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== lanes 1"; DVB_CNN_LANES=1 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== lanes 2"; DVB_CNN_LANES=2 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== lanes 1, old pools"; DVB_CNN_LANES=1 DVB_CNN_POOL_TILED=0 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== lanes 1, old tiles"; DVB_CNN_LANES=1 DVB_CNN_TILE_ANY_N=0 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== halo min 250"; DVB_HALO_MIN_PIXELS=250 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
echo "== chunk 4096"; timeout 200 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3
echo "== chunk 1024"; timeout 200 python tools/cnn_time.py --batch 8192 --chunk 1024 --steps 3
echo "== precise"; timeout 200 python tools/cnn_time.py --batch 2048 --chunk 1024 --steps 3 --precision 1
DVB_HALO_MIN_PIXELS=250 timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q -m gpu -k "block_outputs or branch" 2>&1 | tail -3
