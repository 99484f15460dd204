#!/bin/bash
# Round 2, GPU call 10: hardware-suspended barrier waits, two K blocks per elected round, N-concatenated split MMA, batched shuffles.
mkdir -p gpurun_out
export DVB_TEST_PAIR=1
timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_zz_allele_count_gpu.py -q -m gpu -p no:cacheprovider -k "not allele and not candidates and not run_deepvariant and not make_examples and not device_counts and not long_interval" > gpurun_out/c10_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/c10_pytest.log
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c10_cnn_time.json 2>&1; cat gpurun_out/c10_cnn_time.json
DVB_CNN_PAIR=1 timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c10_cnn_time_pair.json 2>&1; cat gpurun_out/c10_cnn_time_pair.json
timeout 300 python tools/cnn_time.py --batch 4096 --chunk 2048 --steps 3 --warmup 2 --precision 1 > gpurun_out/c10_cnn_time_p1.json 2>&1; cat gpurun_out/c10_cnn_time_p1.json
DVB_CNN_SPLIT_CAT=0 timeout 300 python tools/cnn_time.py --batch 4096 --chunk 2048 --steps 3 --warmup 2 --precision 1 > gpurun_out/c10_cnn_time_p1_nocat.json 2>&1; cat gpurun_out/c10_cnn_time_p1_nocat.json
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 --pacbio > gpurun_out/c10_cnn_time_pacbio.json 2>&1; cat gpurun_out/c10_cnn_time_pacbio.json
DVB_CNN_TRACE=1 timeout 300 python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 0 > /dev/null 2> gpurun_out/c10_trace.err; grep -A10 "rows trace" gpurun_out/c10_trace.err | tail -11
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c10_launches.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu launches exit $?"
DVB_CNN_PAIR=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c10_launches_pair.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1
