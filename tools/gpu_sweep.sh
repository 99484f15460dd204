#!/bin/bash
# one-shot sweep of the generic-kernel pipeline knobs (development aid)
run() { echo "== $*"; env "$@" timeout 200 python tools/cnn_time.py --batch 16384 --chunk 4096 --steps 3 | cut -c1-140; }
run DVB_CNN_SMEM_KB=72
run DVB_CNN_SMEM_KB=56
run DVB_CNN_SMEM_KB=96
run DVB_CNN_SMEM_KB=110 DVB_CNN_MAX_STAGES=4
run DVB_CNN_MAX_STAGES=2
run DVB_CNN_MAX_STAGES=4
run DVB_CNN_LANES=3
run DVB_PERSIST_SMEM_KB=160
echo "== chunk 8192"; timeout 200 python tools/cnn_time.py --batch 16384 --chunk 8192 --steps 3 | cut -c1-140
