#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
N=$1
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --per-config none --chain on --no-cpu-baseline > gpurun_out/c36_bench_n$N.json 2> gpurun_out/c36_bench_n$N.err
echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/c36_bench_n$N.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"]["bit_exact"], d["parity"]["gathered_ok"])
print("chain", d.get("single_chain"))
PY
