#!/bin/bash
# Round 2, GPU call 41: stem_conv1_kernel with __launch_bounds__ minimum blocks 2 / 3 / 4 (shipped); duration of its launches under ncu.
mkdir -p gpurun_out
run() { name=$1; ctas=$2; lib=$3; if [ -n "$lib" ]; then export DVB_LIB_PATH=$PWD/_variants/$lib; else unset DVB_LIB_PATH; fi
  DVB_STEM_CTAS_PER_SM=$ctas timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stem_conv1 -c 4 --csv --log-file gpurun_out/c41_$name.csv python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1
  echo "$name: $(grep stem_conv1 gpurun_out/c41_$name.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"; }
run shipped4 4 ""
run mb3 3 libdvb_stem3.so
run mb2 2 libdvb_stem2.so
run mb3_4ctas 4 libdvb_stem3.so
run shipped4_again 4 ""
