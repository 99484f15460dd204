#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_gpu.py tests/test_encoder_gpu.py tests/test_bam_native.py tests/test_make_examples_native.py -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pipe.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_pipe.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['clocks'])"
DVB_E2E_PIPELINE=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopipe.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_nopipe.json')); print('no pipeline: value', d['value'], 'e2e', d['e2e']['value'])"
