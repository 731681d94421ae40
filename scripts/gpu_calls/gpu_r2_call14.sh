#!/bin/bash
# 8 GPUs: the multi-GPU test, then the scaling bench (cfg2 weak + cfg4 strong 16 GiB + cfg5 64 GiB), reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi topo -m > $O/c14_topo.txt 2>&1; nproc >> $O/c14_topo.txt; free -g >> $O/c14_topo.txt
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > $O/c14_multigpu_test.log 2>&1; echo "pytest rc=$?" >> $O/c14_multigpu_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > $O/c14_bench_n8.json 2> $O/c14_bench_n8.err; echo "rc=$?" >> $O/c14_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > $O/c14_bench_n4.json 2> $O/c14_bench_n4.err; echo "rc=$?" >> $O/c14_bench_n4.err
tail -3 $O/c14_multigpu_test.log; tail -c 400 $O/c14_bench_n8.err; tail -c 300 $O/c14_bench_n4.err
