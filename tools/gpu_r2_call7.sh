#!/bin/bash
# Round 2, GPU call 7: shared-space LDS/STS everywhere (was generic LD/ST), tile allele kernel with hoisted loads.
mkdir -p gpurun_out
export DVB_TEST_PAIR=1
timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_zz_allele_count_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/c7_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/c7_pytest.log
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c7_cnn_time.json 2>&1; cat gpurun_out/c7_cnn_time.json
DVB_CNN_PAIR=1 timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c7_cnn_time_pair.json 2>&1; cat gpurun_out/c7_cnn_time_pair.json
timeout 300 python tools/cnn_time.py --batch 4096 --chunk 2048 --steps 3 --warmup 2 --precision 1 > gpurun_out/c7_cnn_time_p1.json 2>&1; cat gpurun_out/c7_cnn_time_p1.json
DVB_CNN_TRACE=1 timeout 300 python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 0 > /dev/null 2> gpurun_out/c7_trace.err; grep -A12 "rows trace" gpurun_out/c7_trace.err | head -30
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c7_launches.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu launches exit $?"
timeout 600 python tools/allele_count_time.py --mbases 4 > gpurun_out/c7_allele_count_time.json 2> gpurun_out/c7_allele_count_time.err; echo "allele time exit $?"; python -c "
import json; d=json.load(open('gpurun_out/c7_allele_count_time.json')); d.pop('peaks',None); print(d)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c7_launches_allele.csv python tools/allele_count_time.py --mbases 2 --steps 2 --warmup 1 > /dev/null 2>&1; grep -i "allele" gpurun_out/c7_launches_allele.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -4
