#!/bin/bash
# Round 2, 2-GPU: the GAN legs of the bench line with the plain all-reduce default (as the driver's scaling run launches it).
mkdir -p gpurun_out
timeout 55 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-train --no-compress --no-cpu-baseline --no-eager > gpurun_out/c21_bench_2gpu.json 2> gpurun_out/c21_bench_2gpu.err
echo "bench 2gpu rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c21_bench_2gpu.json") if l.startswith("{")][-1])
print("fwd", round(d["value"]), "gan", d.get("gan_generator_iteration_ms"), d.get("gan_discriminator_iteration_ms"), "c3", (d.get("c3_gan_train_iteration") or {}).get("ms_per_generator_iteration"))
PY
