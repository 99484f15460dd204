#!/bin/bash
# Final validation of the round: every GPU test, the bench line, the launch list of the bench command, full-set ncu captures.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --batch 4096 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dvb_encode -s 3 -c 1 -o gpurun_out/enc_full -f python tools/enc_time.py --batch 8192 --steps 2 --warmup 3 > gpurun_out/enc_full.log 2>&1; echo "ncu enc exit $?"
ncu -i gpurun_out/enc_full.ncu-rep --page raw --csv > gpurun_out/enc_full_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:"conv_|pool|stem_|tail" -s 80 -c 80 -o /tmp/cnn_full -f python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > gpurun_out/cnn_full.log 2>&1; echo "ncu cnn exit $?"
ncu -i /tmp/cnn_full.ncu-rep --page raw --csv > gpurun_out/cnn_full_raw.csv 2>/dev/null
du -sh gpurun_out
