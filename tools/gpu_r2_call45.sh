#!/bin/bash
# Round 2, GPU call 45: final check of the committed code - the whole GPU suite, smoke(), the bench line, the encoder alone.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c45_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/c45_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c45_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/c45_smoke.log
timeout 200 python tools/enc_time.py --batch 16384 --steps 20 --warmup 5 > gpurun_out/c45_enc.json 2>/dev/null; cut -c1-140 gpurun_out/c45_enc.json
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/c45_bench_n1.json 2> gpurun_out/c45_bench_err.txt; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/c45_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline_encoder']['frac'], d['roofline_encoder']['achieved'], d['config5_encode_only']['windows_per_s'], d['config5_encode_only']['frac'], d['parity']['images_equal'], d['precision1']['value'], d['clocks'])"
