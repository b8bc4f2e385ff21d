#!/bin/bash
# gradient-exchange A/B at N GPUs (light bench: headline, e2e, sustained loop, strong scaling)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --light > gpurun_out/r2n_bench${N}_peer.json 2> gpurun_out/r2n_bench${N}_peer.err
DUDF_DP_PEER=0 timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --light > gpurun_out/r2n_bench${N}_nccl1.json 2> gpurun_out/r2n_bench${N}_nccl1.err
timeout 200 python bench.py --steps 20 --warmup 3 --light > gpurun_out/r2n_bench1.json 2>gpurun_out/r2n_bench1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"], 4), round(d["value"] / 1e6, 2), d["aux"].get("gradient_exchange", "")[:45], round(d["aux"]["sustained_ms_per_step"], 4), d["aux"].get("strong_scaling_ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2n_bench${N}_peer.err | cut -c1-300
