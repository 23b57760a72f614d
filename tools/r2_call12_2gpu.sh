#!/bin/bash
# Round 2, 2-GPU bench: training step with the in-backward reducer vs the plain after-backward all-reduce (NCCL).
mkdir -p gpurun_out
S=gpurun_out/c12_status.txt
: > $S
N=2
for mode in 1 0; do
  HFC_OVERLAP_ALLREDUCE=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2951$mode bench.py --gpus $N --steps 32 --warmup 3 --no-gan --no-cpu-baseline --no-compress --no-eager \
    > gpurun_out/c12_bench_${N}gpu_overlap$mode.json 2> gpurun_out/c12_bench_${N}gpu_overlap$mode.err
  echo "bench N=$N overlap=$mode rc=$?" >> $S
done
HFC_OVERLAP_ALLREDUCE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-compress --no-eager \
    > gpurun_out/c12_bench_${N}gpu_gan.json 2> gpurun_out/c12_bench_${N}gpu_gan.err; echo "bench gan rc=$?" >> $S
cat $S
python - <<'PY'
import json
for f in ("overlap1", "overlap0", "gan"):
    try:
        lines = [l for l in open(f"gpurun_out/c12_bench_2gpu_{f}.json") if l.startswith("{")]
        d = json.loads(lines[-1]); t = d["train_step"]; g = d.get("gan_train_iteration") or {}
        print(f, "fwd img/s", round(d["value"]), "train ms", round(t["ms_per_step"], 2), "steps", t["steps"], "|", t["gradient_allreduce"][:110],
              "| gan G", g.get("ms_per_generator_iteration"), "D", g.get("ms_per_discriminator_iteration"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
