#!/bin/bash
# Re-entry validation: gpu tests, bench line, launch list of the bench command, ncu full of encoder + CNN kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --batch 4096 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dvb_encode -s 3 -c 1 -o gpurun_out/enc_full -f python tools/enc_time.py --batch 8192 --steps 2 --warmup 3 > gpurun_out/enc_full.log 2>&1; echo "ncu enc exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_" -s 94 -c 94 -o gpurun_out/cnn_full -f python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > gpurun_out/cnn_full.log 2>&1; echo "ncu cnn exit $?"
ls -la gpurun_out
