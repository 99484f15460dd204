#!/bin/bash
# Round 2, GPU call 18 (re-entry): the whole GPU suite on HEAD, then the open A/B switches (halo rule 2, CTA-pair rule 3, stem pipe)
# one by one under tools/cnn_time.py, and the launch lists of precision 0 / precision 1.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c18_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/c18_pytest.log
t() { name=$1; shift; env "$@" timeout 200 python tools/cnn_time.py --batch 8192 --chunk 4096 > gpurun_out/c18_time_$name.json 2> gpurun_out/c18_err_$name.txt; echo "$name exit $?: $(cat gpurun_out/c18_time_$name.json)"; }
t default DVB_NOP=1
t rule2 DVB_HALO_RULE=2
t rule2_t1 DVB_HALO_RULE=2 DVB_HALO_T=1
t pair3 DVB_CNN_PAIR=3 DVB_PERSIST_MIN_N=128
t pair3_tiles2 DVB_CNN_PAIR=3 DVB_PERSIST_MIN_N=128 DVB_PERSIST_MIN_TILES_PER_SM=2
t stem_pipe DVB_STEM_PIPE=1
t all DVB_HALO_RULE=2 DVB_CNN_PAIR=3 DVB_PERSIST_MIN_N=128
timeout 200 python tools/cnn_time.py --batch 4096 --chunk 2048 --precision 1 > gpurun_out/c18_time_p1.json 2>/dev/null; echo "p1 exit $?"; cat gpurun_out/c18_time_p1.json
DVB_HALO_RULE=2 DVB_HALO_SPLIT=1 timeout 200 python tools/cnn_time.py --batch 4096 --chunk 2048 --precision 1 > gpurun_out/c18_time_p1_rule2.json 2>/dev/null; echo "p1 rule2 split exit $?"; cat gpurun_out/c18_time_p1_rule2.json
DVB_CNN_LIST=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c18_launches_p0.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2> gpurun_out/c18_list_p0.txt; echo "ncu p0 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c18_launches_p1.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 --precision 1 > /dev/null 2>&1; echo "ncu p1 exit $?"
timeout 300 python tools/allele_count_time.py > gpurun_out/c18_allele_count_time.json 2> gpurun_out/c18_allele_err.txt; echo "allele exit $?"; tail -c 1500 gpurun_out/c18_allele_count_time.json
