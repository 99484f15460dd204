#!/bin/bash
DVB_CNN_TRACE=1 timeout 300 python tools/cnn_time.py --batch 1024 --chunk 1024 --steps 1 --warmup 0 2>&1 | head -90
