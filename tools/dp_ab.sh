TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 500 $TR tools/dp_check.py > gpurun_out/r2m_dpcheck.txt 2>&1
grep -v "^W1\|^\[W\|Warning\|^\*\|OMP_NUM" gpurun_out/r2m_dpcheck.txt | tail -40
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --light > gpurun_out/r2m_bench2_peer.json 2> gpurun_out/r2m_bench2_peer.err
DUDF_DP_PEER=0 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --light > gpurun_out/r2m_bench2_nccl2.json 2> gpurun_out/r2m_bench2_nccl2.err
DUDF_DP_PEER=0 DUDF_DP_GROUPS=1 timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --light > gpurun_out/r2m_bench2_nccl1.json 2> gpurun_out/r2m_bench2_nccl1.err
timeout 200 python bench.py --steps 20 --warmup 3 --light > gpurun_out/r2m_bench1.json 2>gpurun_out/r2m_bench1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"], 4), round(d["value"] / 1e6, 2), d["aux"].get("gradient_exchange", "")[:45], round(d["aux"]["sustained_ms_per_step"], 4), d["aux"].get("strong_scaling_ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r2m_bench2_peer.err | cut -c1-300
