#!/bin/bash
# Round 2, GPU call 16: the halo kernel under rule 2 (T = 1 / 2 chains, any N block count) on conv5 and the 35x35 3x3 layers:
# plan listing, time against rule 1, parity tests under rule 2, launch lists (precision 0 under both rules, precision 1).
mkdir -p gpurun_out
DVB_CNN_LIST=1 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 4096 > gpurun_out/c16_time_rule1.json 2> gpurun_out/c16_list_rule1.txt; echo "rule1 exit $?"; cat gpurun_out/c16_time_rule1.json
DVB_HALO_RULE=2 DVB_CNN_LIST=1 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 4096 > gpurun_out/c16_time_rule2.json 2> gpurun_out/c16_list_rule2.txt; echo "rule2 exit $?"; cat gpurun_out/c16_time_rule2.json; grep halo gpurun_out/c16_list_rule2.txt; tail -3 gpurun_out/c16_list_rule2.txt | cut -c1-300
DVB_HALO_RULE=2 DVB_HALO_T=1 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 4096 > gpurun_out/c16_time_rule2_t1.json 2>/dev/null; echo "rule2 T1 exit $?"; cat gpurun_out/c16_time_rule2_t1.json
DVB_HALO_RULE=2 timeout 900 python -m pytest tests/test_cnn_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/c16_pytest_rule2.log 2>&1; echo "pytest rule2 exit $?"; tail -5 gpurun_out/c16_pytest_rule2.log
DVB_HALO_RULE=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c16_launches_rule2.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu rule2 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c16_launches_p1.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 --precision 1 > /dev/null 2>&1; echo "ncu p1 exit $?"
DVB_CNN_PAIR=3 DVB_PERSIST_MIN_N=128 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 4096 > gpurun_out/c16_time_pair3.json 2>/dev/null; echo "pair3 exit $?"; cat gpurun_out/c16_time_pair3.json
DVB_CNN_PAD_CIN64_MIN=160 DVB_CNN_PAIR=3 DVB_PERSIST_MIN_N=128 DVB_CNN_LIST=1 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 4096 > gpurun_out/c16_time_pair3_pad.json 2> gpurun_out/c16_list_pair3_pad.txt; echo "pair3+pad exit $?"; cat gpurun_out/c16_time_pair3_pad.json
DVB_HALO_RULE=2 DVB_CNN_PAD_CIN64_MIN=160 DVB_CNN_PAIR=3 DVB_PERSIST_MIN_N=128 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 4096 > gpurun_out/c16_time_all.json 2>/dev/null; echo "all exit $?"; cat gpurun_out/c16_time_all.json
DVB_CNN_PAIR=3 DVB_PERSIST_MIN_N=128 DVB_PERSIST_MIN_TILES_PER_SM=2 timeout 200 python tools/cnn_time.py --batch 8192 --chunk 4096 > gpurun_out/c16_time_pair3_tiles2.json 2>/dev/null; echo "pair3 tiles2 exit $?"; cat gpurun_out/c16_time_pair3_tiles2.json
