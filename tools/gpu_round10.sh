#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/cnn_probe.py 3 > gpurun_out/probe.log 2>&1; echo "probe exit $?"; grep -E "rel_err" gpurun_out/probe.log | awk '{print $6, $1}' | sort -g -r | head -3; tail -2 gpurun_out/probe.log | cut -c1-120
timeout 120 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3
DVB_CNN_HALO=0 timeout 120 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3
DVB_CNN_TRACE=1 timeout 300 python tools/cnn_time.py --batch 1024 --chunk 1024 --steps 1 --warmup 0 2>&1 | grep -A8 "halo trace" | head -30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_|pool3x3|stem_patch|tail" -s 109 -c 109 --csv --log-file gpurun_out/launches_cnn.csv python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > /dev/null 2>&1
