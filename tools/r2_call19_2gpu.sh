#!/bin/bash
# Round 2, 2-GPU: the three gradient all-reduce modes of the training step (plain / pipelined with Adam / in-backward).
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 2 --steps 24 --warmup 5 --no-gan --no-compress --no-cpu-baseline --no-eager > gpurun_out/c19_bench_2gpu.json 2> gpurun_out/c19_bench_2gpu.err
echo "bench 2gpu rc=$?"
grep -v "Warning\|warn\|self._lpips" gpurun_out/c19_bench_2gpu.err | tail -8
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c19_bench_2gpu.json") if l.startswith("{")][-1])
t = d["train_step"]
print("fwd", round(d["value"]), "train", round(t["ms_per_step"], 2), "|", t["gradient_allreduce"][:100], "|", {k: round(v, 2) for k, v in t["phases"].items() if k.endswith("_ms")})
for o in t.get("other_allreduce_modes") or []:
    print("   other:", round(o["ms_per_step"], 2), o["gradient_allreduce"][:80], {k: round(v, 2) for k, v in o["phases"].items() if k.endswith("_ms")})
PY
