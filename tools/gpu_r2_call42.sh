#!/bin/bash
# Round 2, GPU call 42: hot instructions and stall reasons of the final image kernel (128-thread CTAs, word loads).
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dvb_encode_kernel -s 4 -c 1 -o gpurun_out/c42_enc -f python tools/enc_time.py --batch 16384 --steps 2 --warmup 4 > gpurun_out/c42_enc.log 2>&1; echo "ncu exit $?"
python tools/ncu_hot.py gpurun_out/c42_enc.ncu-rep 0 60 > gpurun_out/c42_enc_hot.txt 2>&1; head -3 gpurun_out/c42_enc_hot.txt
ncu -i gpurun_out/c42_enc.ncu-rep --page raw --csv > gpurun_out/c42_enc_raw.csv 2>/dev/null; rm -f gpurun_out/c42_enc.ncu-rep
