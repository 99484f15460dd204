#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_gpu.py tests/test_bam_native.py -x -q -m gpu > gpurun_out/pytest_25.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_25.log
timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
DVB_CNN_MAXPOOL_H2=0 timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pool" -c 26 --csv --log-file gpurun_out/launches_pool.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; grep pool gpurun_out/launches_pool.csv | tail -13 | awk -F'","' '{print $5, $9, $NF}' | cut -c1-120
