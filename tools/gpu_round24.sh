#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_gpu.py -x -q -m gpu > gpurun_out/pytest_24.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_24.log
timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
DVB_STEM_CTAS_PER_SM=3 timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"stem_" -c 2 --csv --log-file gpurun_out/launches_stem.csv python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > /dev/null 2>&1; grep stem_ gpurun_out/launches_stem.csv | awk -F'","' '{print $(NF-2), $NF}'
