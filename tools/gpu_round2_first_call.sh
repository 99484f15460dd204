#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU minutes were spent, cheapest and safest first.
#   1. the new GPU tests alone (allele-count kernels, BAM -> VCF flow, CTA-pair GEMM), each under its own timeout so that a
#      trap / hang in one does not take the others with it;
#   2. allele-count timing against the HBM roofline + its ncu launch list and one full-set capture;
#   3. the classifier with and without DVB_CNN_PAIR=1 (A/B on the same box), and an ncu launch list of the pair run.
mkdir -p gpurun_out
export DVB_TEST_PAIR=1      # the CTA-pair GEMM check is opt-in (experimental kernel, child process under a timeout)
T=tests/test_zz_allele_count_gpu.py
for k in test_make_examples_cli_generates_candidates_on_gpu test_run_deepvariant_from_bam_to_vcf test_device_counts_and_flags \
         test_gpu_candidates_equal_host test_one_launch_over_a_long_interval test_cta_pair_kernel; do
  timeout 300 python -m pytest $T -x -q -m gpu -k $k > gpurun_out/zz_$k.log 2>&1; echo "$k exit $?"; tail -3 gpurun_out/zz_$k.log
done
timeout 600 python tools/allele_count_time.py --mbases 4 > gpurun_out/allele_count_time.json 2> gpurun_out/allele_count_time.err; echo "allele time exit $?"; cat gpurun_out/allele_count_time.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_allele_count.csv python tools/allele_count_time.py --mbases 2 --steps 2 --warmup 1 > /dev/null 2>&1; echo "ncu allele launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dvb_allele_count -s 2 -c 1 -o gpurun_out/allele_full -f python tools/allele_count_time.py --mbases 2 --steps 2 --warmup 1 > gpurun_out/allele_full.log 2>&1; echo "ncu allele full exit $?"
ncu -i gpurun_out/allele_full.ncu-rep --page raw --csv > gpurun_out/allele_full_raw.csv 2>/dev/null
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/cnn_time_default.json 2>&1; cat gpurun_out/cnn_time_default.json
DVB_CNN_PAIR=1 timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/cnn_time_pair.json 2>&1; echo "pair exit $?"; cat gpurun_out/cnn_time_pair.json
DVB_CNN_PAIR=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cnn_pair.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu pair launches exit $?"
du -sh gpurun_out
