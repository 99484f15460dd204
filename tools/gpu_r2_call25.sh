#!/bin/bash
# Round 2, GPU call 25: the ABI-v4 GPU suite once more (multi-sample / aux / stream code is host-side, the encoder changed), and the
# classifier at chunk 8192 against chunk 4096.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c25_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/c25_pytest.log
timeout 600 python tools/cnn_chunk_check.py 16384 > gpurun_out/c25_chunk.json 2> gpurun_out/c25_chunk_err.txt; echo "chunk exit $?"; cat gpurun_out/c25_chunk.json; tail -3 gpurun_out/c25_chunk_err.txt
