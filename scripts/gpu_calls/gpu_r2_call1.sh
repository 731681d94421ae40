#!/bin/bash
# round 2, call 1 (2 GPUs): GPU test suite incl. the 2-rank gather test, bench N=1 (all configs), bench N=2, host-copy probe, engine harness
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/c1_smi.txt 2>&1
nproc > $O/c1_host.txt; free -g >> $O/c1_host.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/c1_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/c1_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/c1_bench_n1.json 2> $O/c1_bench_n1.err; echo "rc=$?" >> $O/c1_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/c1_bench_n2.json 2> $O/c1_bench_n2.err; echo "rc=$?" >> $O/c1_bench_n2.err
timeout 120 scripts/micro/hostcopy > $O/c1_hostcopy.txt 2>&1
MMOORE_PROFILE=1 timeout 300 benchmarks/bench_search 0.2 engine > $O/c1_bench_search_gpu.txt 2> $O/c1_bench_search_gpu.err
timeout 300 oracle/_ref/bench_search_ref 0.2 engine > $O/c1_bench_search_ref.txt 2>&1
tail -3 $O/c1_gpu_tests.log; tail -c 600 $O/c1_bench_n1.err; tail -c 600 $O/c1_bench_n2.err; head -c 1500 $O/c1_bench_n1.json
