#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_35.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_35.log
echo "== encoder (final)"; timeout 300 python tools/enc_time.py --batch 16384 --steps 10
echo "== pacbio"; timeout 300 python tools/enc_time.py --batch 16384 --steps 10 --pacbio
timeout 300 python tools/enc_time.py --batch 16384 --steps 10
