#!/bin/bash
# Round 2, GPU call 6: rows kernel without run-time divisions; which allele kernels run and how long; write-only HBM ceiling.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cnn_gpu.py -q -m gpu -p no:cacheprovider -x -k "rows or stem or all_block" > gpurun_out/c6_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/c6_pytest.log
timeout 300 python tools/cnn_time.py --batch 8192 --chunk 4096 --steps 3 --warmup 2 > gpurun_out/c6_cnn_time.json 2>&1; cat gpurun_out/c6_cnn_time.json
DVB_CNN_TRACE=1 timeout 300 python tools/cnn_time.py --batch 4096 --chunk 4096 --steps 1 --warmup 0 > /dev/null 2> gpurun_out/c6_trace.err; grep -A14 "rows trace" gpurun_out/c6_trace.err | head -34
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c6_launches_allele.csv python tools/allele_count_time.py --mbases 2 --steps 2 --warmup 1 > /dev/null 2>&1; echo "ncu allele exit $?"; grep -i "allele" gpurun_out/c6_launches_allele.csv | awk -F'","' '{print $5, $NF}' | tail -8
python - <<'PY'
import torch, json
out = {}
x = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
for name, fn in (('memset_zero', lambda: x.zero_()), ('fill_7', lambda: x.fill_(7))):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10): fn()
  e1.record(); torch.cuda.synchronize()
  out[name + '_GBps'] = (1 << 30) * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
y = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
out['copy_read_plus_write_GBps'] = 2 * (1 << 30) * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9
print(json.dumps(out))
open('gpurun_out/c6_write_ceiling.json', 'w').write(json.dumps(out))
PY
