#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_abi.py -x -q -m gpu > gpurun_out/pytest_cnn.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_cnn.log
timeout 600 python tools/cnn_precision.py 6 2>&1 | tail -6
timeout 300 python tools/cnn_time.py --batch 2048 --chunk 1024 --steps 3 --precision 1
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 2048 --steps 3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --batch 4096 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dvb_encode -s 3 -c 1 -o gpurun_out/enc_full -f python tools/enc_time.py --batch 8192 --steps 2 --warmup 3 > gpurun_out/enc_full.log 2>&1; echo "ncu enc exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"conv_|pool3x3|stem_patch|tail" -s 109 -c 109 -o /tmp/cnn_full -f python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > gpurun_out/cnn_full.log 2>&1; echo "ncu cnn exit $?"
ncu -i /tmp/cnn_full.ncu-rep --page raw --csv > gpurun_out/cnn_full_raw.csv 2>/dev/null; ls -la /tmp/cnn_full.ncu-rep
du -sh gpurun_out
