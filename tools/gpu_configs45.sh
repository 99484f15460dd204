#!/bin/bash
# BASELINE.json configs 4 (PACBIO layout) and 5 (encode-only sweep) at N = 1
mkdir -p gpurun_out
echo "== config 4: PACBIO 100x147x10 classifier"; timeout 300 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3 --pacbio
echo "== config 4: PACBIO encoder"; timeout 300 python tools/enc_time.py --batch 16384 --steps 10 --pacbio
echo "== config 5: encode-only bench line"; timeout 400 python bench.py --stage encode --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_encode_only.json 2>/dev/null; cat gpurun_out/bench_encode_only.json | cut -c1-900
