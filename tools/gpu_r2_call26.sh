#!/bin/bash
# Round 2, GPU call 26: CUDA-graph replay of the chunk forward (DVB_CNN_GRAPH=1) - time against direct launches on a side stream,
# parity suite under the switch, both precisions.
mkdir -p gpurun_out
t() { name=$1; shift; env "$@" timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 6 --warmup 3 > gpurun_out/c26_time_$name.json 2> gpurun_out/c26_err_$name.txt; echo "$name exit $?: $(cat gpurun_out/c26_time_$name.json)"; tail -2 gpurun_out/c26_err_$name.txt; }
t default DVB_NOP=1
t side_stream DVB_CNN_SIDE_STREAM=1
t graph DVB_CNN_GRAPH=1
t graph_again DVB_CNN_GRAPH=1
t default_again DVB_NOP=1
DVB_CNN_GRAPH=1 timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_fused.py tests/test_call_variants.py -m gpu -x -q -p no:cacheprovider > gpurun_out/c26_pytest_graph.log 2>&1; echo "pytest graph exit $?"; tail -3 gpurun_out/c26_pytest_graph.log
DVB_CNN_GRAPH=1 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --precision 1 > gpurun_out/c26_time_p1_graph.json 2>/dev/null; echo "p1 graph: $(cat gpurun_out/c26_time_p1_graph.json)"
DVB_CNN_SIDE_STREAM=1 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 2048 --precision 1 > gpurun_out/c26_time_p1.json 2>/dev/null; echo "p1: $(cat gpurun_out/c26_time_p1.json)"
