#!/bin/bash
# Round 2, GPU call 32: encoder A/B - DVB_ENC_MIN_BLOCKS 2 with and without the word loads of the 4-pixel groups, and the encoder /
# golden / KAT suites on the word-load build.
mkdir -p gpurun_out
for rep in 1 2; do
for v in mb2w mb2s mb2sw; do
  if [ $v = shipped ]; then unset DVB_LIB_PATH; else export DVB_LIB_PATH=$PWD/_variants/libdvb_$v.so; fi
  timeout 200 python tools/enc_time.py --batch 16384 --steps 20 --warmup 5 > gpurun_out/c32_enc_${v}_$rep.json 2>/dev/null; echo "$v $rep: $(cut -c1-130 gpurun_out/c32_enc_${v}_$rep.json)"
done; done
export DVB_LIB_PATH=$PWD/_variants/libdvb_mb2sw.so
timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_golden.py tests/test_pileup_kat.py tests/test_channel_planes.py tests/test_pair_support.py -m gpu -x -q -p no:cacheprovider > gpurun_out/c32_pytest_mb2w.log 2>&1; echo "pytest mb2sw exit $?"; tail -2 gpurun_out/c32_pytest_mb2w.log
