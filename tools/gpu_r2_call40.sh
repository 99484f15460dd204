#!/bin/bash
# Round 2, GPU call 40: encoder CTA size - 128 / 256 (shipped) / 512 threads per image CTA.
mkdir -p gpurun_out
for rep in 1 2; do for v in shipped t96 t192 p32 p64 p256; do
  if [ $v = shipped ]; then unset DVB_LIB_PATH; else export DVB_LIB_PATH=$PWD/_variants/libdvb_$v.so; fi
  timeout 200 python tools/enc_time.py --batch 16384 --steps 20 --warmup 5 > gpurun_out/c40_enc_${v}_$rep.json 2>/dev/null; echo "$v $rep: $(cut -c1-130 gpurun_out/c40_enc_${v}_$rep.json)"
done; done
