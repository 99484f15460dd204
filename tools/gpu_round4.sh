#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section SpeedOfLight_RooflineChart --clock-control none -k regex:conv_gemm -s 94 -c 94 -o /tmp/cnn_r1 python tools/cnn_time.py --batch 1024 --chunk 1024 --steps 1 --warmup 1 > gpurun_out/ncu_cnn.log 2>&1
tail -2 gpurun_out/ncu_cnn.log
ncu -i /tmp/cnn_r1.ncu-rep --page raw --csv > gpurun_out/cnn_r1_raw.csv 2>/dev/null
ls -la /tmp/cnn_r1.ncu-rep gpurun_out/cnn_r1_raw.csv
