#!/bin/bash
# Round 2, GPU call 9: ncu --set full with source counters on the rows kernel (conv3 + pool), one persistent GEMM layer and one
# one-tile-per-CTA layer: where do the warps stall?
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_rows_kernel -s 2 -c 2 -o gpurun_out/c9_rows -f python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > gpurun_out/c9_rows.log 2>&1; echo "ncu rows exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_persistent_kernel -s 40 -c 1 -o gpurun_out/c9_persist -f python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > gpurun_out/c9_persist.log 2>&1; echo "ncu persist exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_kernel -s 50 -c 3 -o gpurun_out/c9_gemm -f python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > gpurun_out/c9_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:stem_conv1_kernel -s 1 -c 1 -o gpurun_out/c9_stem -f python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > gpurun_out/c9_stem.log 2>&1; echo "ncu stem exit $?"
ls -la gpurun_out/*.ncu-rep
