#!/bin/bash
# Round 2, 2-GPU check of the bench line as the driver's scaling run launches it (both gradient all-reduce modes timed).
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c18_bench_2gpu.json 2> gpurun_out/c18_bench_2gpu.err
echo "bench 2gpu rc=$?"
tail -3 gpurun_out/c18_bench_2gpu.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c18_bench_2gpu.json") if l.startswith("{")][-1])
t = d["train_step"]
print("fwd", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "train", t["ms_per_step"], "|", t["gradient_allreduce"][:90])
print("other:", json.dumps(t.get("other_allreduce_mode"))[:400])
print("gan", d.get("gan_generator_iteration_ms"), d.get("gan_discriminator_iteration_ms"))
PY
