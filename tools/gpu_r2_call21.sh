#!/bin/bash
# Round 2, GPU call 21: run-to-run determinism of the classifier plans (a race in the pair kernel under rule 3 was seen once in call 20).
mkdir -p gpurun_out
timeout 600 python tools/cnn_race_check.py 150 6 > gpurun_out/c21_race_b6.txt 2>&1; echo "b6 exit $?"; tail -5 gpurun_out/c21_race_b6.txt
timeout 600 python tools/cnn_race_check.py 40 64 > gpurun_out/c21_race_b64.txt 2>&1; echo "b64 exit $?"; tail -5 gpurun_out/c21_race_b64.txt
