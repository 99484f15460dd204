#!/bin/bash
# Round 2, GPU call 1: the whole GPU suite (pair-kernel check on), the allele-count kernels timed + profiled, the classifier with and
# without the CTA-pair kernel, and the new bench line (parity / precision1 / config4 / config5 blocks).
mkdir -p gpurun_out
export DVB_TEST_PAIR=1
timeout 1200 python -m pytest tests -q -m gpu --durations=12 -p no:cacheprovider > gpurun_out/c1_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/c1_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/c1_smoke.log
timeout 600 python tools/allele_count_time.py --mbases 4 > gpurun_out/c1_allele_count_time.json 2> gpurun_out/c1_allele_count_time.err; echo "allele time exit $?"; cat gpurun_out/c1_allele_count_time.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dvb_allele -s 4 -c 2 -o gpurun_out/c1_allele_full -f python tools/allele_count_time.py --mbases 2 --steps 2 --warmup 1 > gpurun_out/c1_allele_full.log 2>&1; echo "ncu allele full exit $?"
ncu -i gpurun_out/c1_allele_full.ncu-rep --page raw --csv > gpurun_out/c1_allele_full_raw.csv 2>/dev/null
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c1_cnn_time_default.json 2>&1; cat gpurun_out/c1_cnn_time_default.json
DVB_CNN_PAIR=1 timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c1_cnn_time_pair.json 2>&1; echo "pair exit $?"; tail -3 gpurun_out/c1_cnn_time_pair.json
timeout 300 python tools/cnn_time.py --batch 4096 --chunk 2048 --steps 3 --warmup 2 --precision 1 > gpurun_out/c1_cnn_time_p1.json 2>&1; cat gpurun_out/c1_cnn_time_p1.json
DVB_CNN_PAIR=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c1_launches_cnn_pair.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu pair launches exit $?"
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench exit $?"; cat gpurun_out/c1_bench.json; tail -5 gpurun_out/c1_bench.err
rm -f gpurun_out/c1_allele_full.ncu-rep.tmp; du -sh gpurun_out
