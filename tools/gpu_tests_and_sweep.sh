#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_final.log
echo "== encoder"; timeout 300 python tools/enc_time.py --batch 16384 --steps 10
bash tools/gpu_sweep.sh
