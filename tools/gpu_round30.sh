#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_golden.py tests/test_pileup_kat.py tests/test_make_examples_native.py -x -q -m gpu > gpurun_out/pytest_30.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_30.log
echo "== prepass"; timeout 300 python tools/enc_time.py --batch 16384 --steps 10
echo "== single kernel"; DVB_ENC_PREPASS=0 timeout 300 python tools/enc_time.py --batch 16384 --steps 10
echo "== prepass pacbio"; timeout 300 python tools/enc_time.py --batch 16384 --steps 10 --pacbio
echo "== single pacbio"; DVB_ENC_PREPASS=0 timeout 300 python tools/enc_time.py --batch 16384 --steps 10 --pacbio
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"dvb_" -s 6 -c 4 --csv --log-file gpurun_out/launches_enc.csv python tools/enc_time.py --batch 8192 --steps 2 --warmup 3 > /dev/null 2>&1; grep "dvb_" gpurun_out/launches_enc.csv | awk -F'","' '{n=split($5,a,"::"); print substr(a[n],1,28), $(NF-2), $NF}'
