#!/bin/bash
# Round 2, GPU call 47: the GPU suite + smoke() after the importer / N-scoring change (device Smith-Waterman kernel rebuilt).
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c47_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/c47_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c47_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/c47_smoke.log
