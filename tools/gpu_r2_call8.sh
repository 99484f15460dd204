#!/bin/bash
# Round 2, GPU call 8: rows kernel with register-direct stores / shuffle pooling; allele tile kernel at 2 x 16 warps per SM;
# conv5 experiments (pad 80 -> 128 channels; all-persistent + CTA pairs); precision-1 launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_gpu.py tests/test_zz_allele_count_gpu.py tests/test_encoder_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/c8_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/c8_pytest.log
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c8_cnn_time.json 2>&1; cat gpurun_out/c8_cnn_time.json
DVB_CNN_TRACE=1 timeout 300 python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 0 > /dev/null 2> gpurun_out/c8_trace.err; grep -A10 "rows trace" gpurun_out/c8_trace.err | head -24
timeout 600 python tools/allele_count_time.py --mbases 4 > gpurun_out/c8_allele_count_time.json 2> gpurun_out/c8_allele_count_time.err; echo "allele time exit $?"; python -c "
import json; d=json.load(open('gpurun_out/c8_allele_count_time.json')); d.pop('peaks',None); print(d)"
timeout 300 python tools/enc_time.py --batch 16384 > gpurun_out/c8_enc_time.json 2>&1; cat gpurun_out/c8_enc_time.json
for v in "DVB_CNN_PAD_CIN64_MIN=80" "DVB_CNN_PERSIST=2" "DVB_CNN_PERSIST=2 DVB_CNN_PAIR=1"; do
  tag=$(echo $v | tr ' =' '__')
  env $v timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c8_cnn_time_$tag.json 2>&1; echo "$v: $(cat gpurun_out/c8_cnn_time_$tag.json)"
  env $v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c8_launches_$tag.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c8_launches.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c8_launches_p1.csv python tools/cnn_time.py --batch 2048 --chunk 2048 --steps 1 --warmup 1 --precision 1 > /dev/null 2>&1; echo "ncu p1 exit $?"
