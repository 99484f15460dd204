#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_call_variants.py -x -q -m gpu > gpurun_out/pytest_22.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_22.log
echo "== default (twin, stem cp.async ring)"; timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
echo "== no twin"; DVB_CNN_TWIN=0 timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
echo "== twin, chunk 2048"; timeout 200 python tools/cnn_time.py --batch 16384 --chunk 2048 --steps 3
echo "== twin, lanes 2"; DVB_CNN_LANES=2 timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
echo "== no twin, stem ctas 5"; DVB_CNN_TWIN=0 DVB_STEM_CTAS_PER_SM=5 timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stem_" -c 4 --csv --log-file gpurun_out/launches_stem.csv python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > /dev/null 2>&1; grep stem_ gpurun_out/launches_stem.csv | cut -d, -f15 | head -4
