#!/bin/bash
# Round 2, GPU call 27: end-of-round evidence on the final code - smoke(), the launch list of the bench command, full-set ncu of one
# classifier forward (2048 images) and of the encoder launch pair.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c27_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/c27_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/c27_launches_bench.csv python bench.py --steps 2 --warmup 1 --batch 4096 --no-cpu-baseline --no-e2e > gpurun_out/c27_bench_under_ncu.log 2>&1; echo "ncu bench launches exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"conv_|pool3x3|stem_|tail" -s 79 -c 79 -o gpurun_out/c27_cnn_full -f python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 > gpurun_out/c27_cnn_full.log 2>&1; echo "ncu full exit $?"
ncu -i gpurun_out/c27_cnn_full.ncu-rep --page raw --csv > gpurun_out/c27_cnn_full_raw.csv 2>/dev/null; rm -f gpurun_out/c27_cnn_full.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:dvb_ -s 6 -c 2 -o gpurun_out/c27_enc_full -f python tools/enc_time.py --batch 8192 --steps 2 --warmup 3 > gpurun_out/c27_enc_full.log 2>&1; echo "ncu enc exit $?"
ncu -i gpurun_out/c27_enc_full.ncu-rep --page raw --csv > gpurun_out/c27_enc_full_raw.csv 2>/dev/null; rm -f gpurun_out/c27_enc_full.ncu-rep
