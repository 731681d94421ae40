#!/bin/bash
# 2 GPUs: NCCL tests (gather, spill, chain search) + headline and single_chain at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu > gpurun_out/c26_multigpu.log 2>&1
echo "rc=$?" >> gpurun_out/c26_multigpu.log
tail -15 gpurun_out/c26_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --per-config none --chain on --no-cpu-baseline > gpurun_out/c26_bench_n2.json 2> gpurun_out/c26_bench_n2.err
echo "bench rc=$?"
tail -c 2500 gpurun_out/c26_bench_n2.json; tail -5 gpurun_out/c26_bench_n2.err
