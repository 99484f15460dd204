#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_golden.py tests/test_pileup_kat.py tests/test_make_examples_native.py tests/test_bam_native.py -x -q -m gpu > gpurun_out/pytest_35.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_35.log
echo "== flat blank tail"; timeout 300 python tools/enc_time.py --batch 16384 --steps 10
echo "== pacbio"; timeout 300 python tools/enc_time.py --batch 16384 --steps 10 --pacbio
timeout 300 python tools/enc_time.py --batch 16384 --steps 10
