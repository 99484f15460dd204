#!/bin/bash
# compute-sanitizer memcheck over one small pass of every kernel (smoke: encoder pre-pass + image kernel, classifier at both precisions)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|smoke\]|Error|error" gpurun_out/sanitizer_smoke.log | head -12
timeout 600 python -m pytest tests/test_golden.py tests/test_bam_native.py tests/test_encoder_gpu.py -x -q -m gpu 2>&1 | tail -3
