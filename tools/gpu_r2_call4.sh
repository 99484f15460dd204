#!/bin/bash
# Round 2, GPU call 4: elect.sync issue (no waterfall loops around UTCHMMA / UTMALDG): parity, speed, microbenchmark, timeline.
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate tools/micro/mma_rate.cu && timeout 120 /tmp/mma_rate > gpurun_out/c4_mma_rate.txt 2>&1; echo "mma_rate exit $?"; cat gpurun_out/c4_mma_rate.txt
export DVB_TEST_PAIR=1
timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_zz_allele_count_gpu.py -q -m gpu -p no:cacheprovider -k "not allele and not candidates and not run_deepvariant and not make_examples" > gpurun_out/c4_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/c4_pytest.log
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c4_cnn_time.json 2>&1; cat gpurun_out/c4_cnn_time.json
DVB_CNN_PAIR=1 timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c4_cnn_time_pair.json 2>&1; cat gpurun_out/c4_cnn_time_pair.json
timeout 300 python tools/cnn_time.py --batch 4096 --chunk 2048 --steps 3 --warmup 2 --precision 1 > gpurun_out/c4_cnn_time_p1.json 2>&1; cat gpurun_out/c4_cnn_time_p1.json
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 --pacbio > gpurun_out/c4_cnn_time_pacbio.json 2>&1; cat gpurun_out/c4_cnn_time_pacbio.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c4_launches.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu launches exit $?"
DVB_CNN_PAIR=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c4_launches_pair.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu pair launches exit $?"
DVB_CNN_TRACE=1 timeout 300 python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 0 > gpurun_out/c4_trace.json 2> gpurun_out/c4_trace.err; grep -A26 "rows trace" gpurun_out/c4_trace.err | head -60
