#!/bin/bash
# Round 2, GPU call 12: stem_rows_kernel (bias alignment fixed), pooled row through shared memory again, batched Smith-Waterman.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ssw_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/c12_pytest_ssw.log 2>&1; echo "pytest ssw exit $?"; tail -12 gpurun_out/c12_pytest_ssw.log
timeout 900 python -m pytest tests/test_cnn_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/c12_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/c12_pytest.log
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c12_cnn_time.json 2>&1; tail -1 gpurun_out/c12_cnn_time.json
DVB_CNN_STEM_ROWS=0 timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c12_cnn_time_oldstem.json 2>&1; tail -1 gpurun_out/c12_cnn_time_oldstem.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c12_launches.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu launches exit $?"
