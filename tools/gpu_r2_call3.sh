#!/bin/bash
# Round 2, GPU call 3: tcgen05.mma issue-rate microbenchmark + timeline of the row-streaming kernel.
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/micro/mma_rate.cu && timeout 120 /tmp/mma_rate > gpurun_out/c3_mma_rate.txt 2>&1; echo "mma_rate exit $?"; cat gpurun_out/c3_mma_rate.txt
DVB_CNN_TRACE=1 timeout 300 python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 0 > gpurun_out/c3_trace.json 2> gpurun_out/c3_trace.err; grep -A42 "rows trace" gpurun_out/c3_trace.err | head -100
