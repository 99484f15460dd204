#!/bin/bash
# Round 2, GPU call 20: the whole GPU suite on the ABI-v4 code (channel planes, blank mask, multi-sample golden, allele-frequency golden).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c20_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/c20_pytest.log
