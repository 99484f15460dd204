#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_golden.py tests/test_pileup_kat.py -x -q -m gpu > gpurun_out/pytest_31.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_31.log
echo "== prepass"; timeout 300 python tools/enc_time.py --batch 16384 --steps 10
echo "== pacbio"; timeout 300 python tools/enc_time.py --batch 16384 --steps 10 --pacbio
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json | cut -c1-700; tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --batch 4096 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dvb_ -s 6 -c 2 -o gpurun_out/enc_full -f python tools/enc_time.py --batch 8192 --steps 2 --warmup 3 > gpurun_out/enc_full.log 2>&1; echo "ncu enc exit $?"
ncu -i gpurun_out/enc_full.ncu-rep --page raw --csv > gpurun_out/enc_full_raw.csv 2>/dev/null
