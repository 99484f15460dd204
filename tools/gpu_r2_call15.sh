#!/bin/bash
# Round 2, GPU call 15 (1 GPU): allele tile kernel with per-lane header fetch (tests + timing), the product CLI's wall clock on config 1
# (quick-start reads) and a synthetic 30x BAM.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_pair_support.py tests/test_zz_allele_count_gpu.py tests/test_ssw_gpu.py tests/test_fused.py tests/test_bam_native.py tests/test_encoder_gpu.py tests/test_cnn_gpu.py -m gpu -q --timeout 300 > gpurun_out/c15_pytest.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/c15_pytest.log | cut -c1-400
timeout 300 python tools/allele_count_time.py > gpurun_out/c15_allele_count_time.json 2> gpurun_out/c15_allele.err; echo "allele time exit $?"; cut -c1-1500 gpurun_out/c15_allele_count_time.json; tail -3 gpurun_out/c15_allele.err
timeout 900 python tools/cli_throughput.py --mbases 0.2 --out gpurun_out/c15_cli_throughput.json > gpurun_out/c15_cli.log 2> gpurun_out/c15_cli.err; echo "cli exit $?"; tail -2 gpurun_out/c15_cli.log | cut -c1-1500; tail -5 gpurun_out/c15_cli.err
NC=$(python -c 'import os; print(min(16, len(os.sched_getaffinity(0))))'); timeout 900 python tools/cli_throughput.py --mbases 0.4 --shards $NC --out gpurun_out/c15_cli_throughput_sharded.json > gpurun_out/c15_cli_sharded.log 2> gpurun_out/c15_cli_sharded.err; echo "cli sharded exit $?"; tail -1 gpurun_out/c15_cli_sharded.log | cut -c1-1500; tail -5 gpurun_out/c15_cli_sharded.err
timeout 900 python -X importtime -c "import deepvariant_b200.cli" 2> gpurun_out/c15_importtime.txt; sort -t'|' -k2 -n gpurun_out/c15_importtime.txt | tail -5
