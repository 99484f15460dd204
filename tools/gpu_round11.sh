#!/bin/bash
for m in 1 2; do echo "== HALO_MODE=$m"; DVB_HALO_MODE=$m timeout 300 python tools/cnn_probe.py 3 2>&1 | grep -E "^s[2-5] |^p1|max \|dp|batch 512" | cut -c1-100; done
