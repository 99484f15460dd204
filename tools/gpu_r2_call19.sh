#!/bin/bash
# Round 2, GPU call 19: the plane-backed channels (ABI v4) on the device - their own tests, the encoder suite, the KATs, the golden
# pins; the encoder's time on the bench workload (no regression from the added instantiations).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_channel_planes.py tests/test_encoder_gpu.py tests/test_pileup_kat.py tests/test_golden.py tests/test_pair_support.py tests/test_abi.py -m gpu -x -q -p no:cacheprovider > gpurun_out/c19_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/c19_pytest.log
timeout 300 python tools/enc_time.py --batch 16384 --steps 10 --warmup 3 > gpurun_out/c19_enc_time.json 2> gpurun_out/c19_enc_err.txt; echo "enc_time exit $?"; cat gpurun_out/c19_enc_time.json
