#!/bin/bash
# Round 2, GPU call 5: burst microbenchmark (hardware MMA interval), timelines of the rows kernel's store phase and of a persistent GEMM
# layer, fused-flow test, tile allele-count kernel (parity + time).
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/micro/mma_rate.cu && timeout 120 /tmp/mma_rate > gpurun_out/c5_mma_rate.txt 2>&1; echo "mma_rate exit $?"; grep "^burst\|^issue" gpurun_out/c5_mma_rate.txt
DVB_CNN_LIST=1 timeout 120 python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 0 2> gpurun_out/c5_list.err > /dev/null
L=$(grep "1x7 .*N=192 .*persist=1" gpurun_out/c5_list.err | head -1 | sed 's/\[conv \([0-9]*\)\].*/\1/'); echo "trace layer $L"; grep "persist=1" gpurun_out/c5_list.err | head -30
DVB_CNN_TRACE=2 DVB_CNN_TRACE_LAYER=$L timeout 300 python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 0 > /dev/null 2> gpurun_out/c5_gemm_trace.err; grep -A70 "gemm trace" gpurun_out/c5_gemm_trace.err | head -75
DVB_CNN_TRACE=1 timeout 300 python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 0 > /dev/null 2> gpurun_out/c5_trace.err; grep -A22 "rows trace" gpurun_out/c5_trace.err | head -50
timeout 600 python -m pytest tests/test_fused.py tests/test_zz_allele_count_gpu.py tests/test_cnn_gpu.py -q -m gpu -p no:cacheprovider -x -k "fused or allele or candidates or device_counts or long_interval or rows" > gpurun_out/c5_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/c5_pytest.log
timeout 600 python tools/allele_count_time.py --mbases 4 > gpurun_out/c5_allele_count_time.json 2> gpurun_out/c5_allele_count_time.err; echo "allele time exit $?"; python -c "
import json; d=json.load(open('gpurun_out/c5_allele_count_time.json')); d.pop('peaks',None); print(d)"
