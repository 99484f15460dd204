#!/bin/bash
# Round 2, GPU call 48: the GPU suite + smoke() on the final tree (derived read tables, importer, CRAM input, stream hardening).
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c48_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/c48_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c48_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/c48_smoke.log
