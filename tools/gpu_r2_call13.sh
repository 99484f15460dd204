#!/bin/bash
# Round 2, GPU call 13: checkpoint - the whole GPU suite, smoke, the bench line, launch list + full-set ncu of one forward.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 -p no:cacheprovider > gpurun_out/c13_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -16 gpurun_out/c13_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c13_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/c13_smoke.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; echo "bench exit $?"; cat gpurun_out/c13_bench.json; tail -3 gpurun_out/c13_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c13_launches.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"conv_|pool3x3|stem_|tail" -s 79 -c 79 -o gpurun_out/c13_cnn_full -f python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > gpurun_out/c13_cnn_full.log 2>&1; echo "ncu full exit $?"
ncu -i gpurun_out/c13_cnn_full.ncu-rep --page raw --csv > gpurun_out/c13_cnn_full_raw.csv 2>/dev/null; rm -f gpurun_out/c13_cnn_full.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:dvb_ -s 6 -c 2 -o gpurun_out/c13_enc_full -f python tools/enc_time.py --batch 8192 --steps 2 --warmup 3 > gpurun_out/c13_enc_full.log 2>&1; echo "ncu enc exit $?"
ncu -i gpurun_out/c13_enc_full.ncu-rep --page raw --csv > gpurun_out/c13_enc_full_raw.csv 2>/dev/null; rm -f gpurun_out/c13_enc_full.ncu-rep
