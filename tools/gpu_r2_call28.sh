#!/bin/bash
# Round 2, GPU call 28 (2 GPUs): the driver-shaped N = 2 line of the end-of-round code (torchrun, NCCL for the barrier and the max time
# only) and the product CLI with two tasks on two GPUs (fused flow) against the staged flow.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/c28_bench_n2.json 2> gpurun_out/c28_bench_n2.err; echo "bench n2 exit $?"; cut -c1-400 gpurun_out/c28_bench_n2.json; tail -2 gpurun_out/c28_bench_n2.err
timeout 600 python -m pytest tests/test_fused.py tests/test_sharding.py -m gpu -q -p no:cacheprovider > gpurun_out/c28_fused.log 2>&1; echo "fused tests exit $?"; tail -2 gpurun_out/c28_fused.log
