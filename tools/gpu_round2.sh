#!/bin/bash
mkdir -p gpurun_out
for mode in 0 2; do
  for bk in 64 0; do
    echo "=== STRIDE_MODE=$mode BLOCK_K=$bk" 
    DVB_TMA_STRIDE_MODE=$mode DVB_CNN_BLOCK_K=$bk timeout 180 python tools/cnn_probe.py 3 > gpurun_out/probe_m${mode}_k${bk}.log 2>&1
    echo "exit $?" >> gpurun_out/probe_m${mode}_k${bk}.log
    tail -45 gpurun_out/probe_m${mode}_k${bk}.log | cut -c1-160
  done
done
