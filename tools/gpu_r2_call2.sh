#!/bin/bash
# Round 2, GPU call 2: the row-streaming stem kernels (parity first), then where the GEMM kernels' time goes (probes + cuBLAS reference).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cnn_gpu.py -q -m gpu -p no:cacheprovider -k "rows or unfused or stem or all_block or pacbio_geometry or bench_scale" > gpurun_out/c2_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/c2_pytest.log
for rows in 1 0; do
  DVB_CNN_ROWS=$rows timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c2_cnn_time_rows$rows.json 2>&1; cat gpurun_out/c2_cnn_time_rows$rows.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c2_launches_rows.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu launches exit $?"
for dbg in 1 2; do
  DVB_CNN_DBG=$dbg timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c2_launches_dbg$dbg.csv python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu dbg$dbg exit $?"
done
timeout 300 python tools/gemm_ref.py 4096 > gpurun_out/c2_gemm_ref.json 2>&1; cat gpurun_out/c2_gemm_ref.json
du -sh gpurun_out
