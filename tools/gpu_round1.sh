#!/bin/bash
# First GPU visit: parity tests, smoke, encoder timing, ncu evidence.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 300 python tools/enc_time.py --batch 16384 --steps 10 > gpurun_out/enc_time_wgs.json 2> gpurun_out/enc_time_wgs.err; cat gpurun_out/enc_time_wgs.json; tail -3 gpurun_out/enc_time_wgs.err
timeout 300 python tools/enc_time.py --batch 16384 --steps 10 --pacbio > gpurun_out/enc_time_pacbio.json 2>> gpurun_out/enc_time_wgs.err; cat gpurun_out/enc_time_pacbio.json
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu -k "odd_shapes or unaligned or empty" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer exit $?" >> gpurun_out/sanitizer.log
tail -5 gpurun_out/sanitizer.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dvb_encode -s 3 -c 2 -o gpurun_out/enc_r1 python tools/enc_time.py --batch 8192 --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
