#!/bin/bash
# Round 2, GPU call 29: the encoder with DVB_ENC_MIN_BLOCKS 3 / 4 (shipped) / 5 (prebuilt variants of the same sources), interleaved.
mkdir -p gpurun_out
for rep in 1 2; do
for v in shipped mb3 mb5; do
  if [ $v = shipped ]; then unset DVB_LIB_PATH; else export DVB_LIB_PATH=$PWD/_variants/libdvb_$v.so; fi
  timeout 200 python tools/enc_time.py --batch 16384 --steps 20 --warmup 5 > gpurun_out/c29_enc_${v}_$rep.json 2>/dev/null; echo "$v $rep: $(cat gpurun_out/c29_enc_${v}_$rep.json)"
done; done
