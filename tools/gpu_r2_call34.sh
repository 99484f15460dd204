#!/bin/bash
# Round 2, GPU call 34: the grid-stride pre-pass (mb2wp build; DVB_ENC_PREPASS_PERSIST=0 restores one CTA per 4 images) against the
# build without the loop (mb2w), and the suites on it.
mkdir -p gpurun_out
for rep in 1 2; do
  DVB_LIB_PATH=$PWD/_variants/libdvb_mb2w.so timeout 200 python tools/enc_time.py --batch 16384 --steps 20 --warmup 5 > gpurun_out/c34_mb2w_$rep.json 2>/dev/null; echo "mb2w $rep: $(cut -c1-130 gpurun_out/c34_mb2w_$rep.json)"
  DVB_LIB_PATH=$PWD/_variants/libdvb_mb2wp.so timeout 200 python tools/enc_time.py --batch 16384 --steps 20 --warmup 5 > gpurun_out/c34_mb2wp_$rep.json 2>/dev/null; echo "mb2wp persist $rep: $(cut -c1-130 gpurun_out/c34_mb2wp_$rep.json)"
  DVB_ENC_PREPASS_PERSIST=0 DVB_LIB_PATH=$PWD/_variants/libdvb_mb2wp.so timeout 200 python tools/enc_time.py --batch 16384 --steps 20 --warmup 5 > gpurun_out/c34_mb2wp0_$rep.json 2>/dev/null; echo "mb2wp one-shot $rep: $(cut -c1-130 gpurun_out/c34_mb2wp0_$rep.json)"
done
export DVB_LIB_PATH=$PWD/_variants/libdvb_mb2wp.so
timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_golden.py tests/test_pileup_kat.py tests/test_channel_planes.py tests/test_pair_support.py -m gpu -x -q -p no:cacheprovider > gpurun_out/c34_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/c34_pytest.log
