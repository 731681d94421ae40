#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "one_chain" > gpurun_out/c25_chainbig.log 2>&1
echo "rc=$?" >> gpurun_out/c25_chainbig.log
tail -25 gpurun_out/c25_chainbig.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "chain or fuzz_large or big_blocks" > gpurun_out/c25_parity.log 2>&1
echo "rc=$?" >> gpurun_out/c25_parity.log
tail -5 gpurun_out/c25_parity.log
timeout 900 python bench.py --per-config none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c25_bench_head.json 2> gpurun_out/c25_bench_head.err
tail -c 1500 gpurun_out/c25_bench_head.json; tail -5 gpurun_out/c25_bench_head.err
python - <<'PY' > gpurun_out/c25_chain_n1.json 2> gpurun_out/c25_chain_n1.err
import json, os, sys, types
sys.path.insert(0, os.getcwd())
import torch, bench
import monkey_moore_b200 as mm
ctx = bench.Ctx(); ctx.torch, ctx.mm = torch, mm
ctx.rank, ctx.world, ctx.local = 0, 1, 0
torch.cuda.set_device(0); ctx.dist = ctx.comm = None; ctx.stream = torch.cuda.current_stream()
args = types.SimpleNamespace(size_mib=0, no_verify=False)
print(json.dumps(bench.measure_chain(ctx, args)))
PY
cat gpurun_out/c25_chain_n1.json; tail -5 gpurun_out/c25_chain_n1.err
