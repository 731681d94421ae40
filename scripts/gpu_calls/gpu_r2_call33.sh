#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/c33_bench_n8.json 2> gpurun_out/c33_bench_n8.err
echo "rc=$?"
tail -c 600 gpurun_out/c33_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c33_bench_n8.json").read().strip().splitlines()[-1])
    print("headline", d["value"], d["ms_per_step"], d["parity"])
    for p in d.get("per_config", []):
        print(p.get("workload"), p.get("value"), p.get("ms_per_step"), (p.get("parity") or {}).get("bit_exact"), (p.get("parity") or {}).get("gathered_ok"), p.get("error"))
    print("chain", d.get("single_chain"))
except Exception as e:
    print("no line:", e)
PY
