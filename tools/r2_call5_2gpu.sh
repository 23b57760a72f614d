#!/bin/bash
# Round 2, 2-GPU call: the in-backward gradient reducer over NCCL (self-check + step time against the plain all-reduce) and
# the train_ddp launcher.   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_call5_2gpu.sh'
mkdir -p gpurun_out
S=gpurun_out/c5_status.txt
: > $S
N=${1:-2}
for mode in 1 0; do
  HFC_OVERLAP_ALLREDUCE=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 2951$mode bench.py --gpus $N --steps 12 --warmup 3 --no-cpu-baseline --no-compress --no-eager \
    > gpurun_out/c5_bench_${N}gpu_overlap$mode.json 2> gpurun_out/c5_bench_${N}gpu_overlap$mode.err
  echo "bench N=$N overlap=$mode rc=$?" >> $S
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  -m hific_b200.train_ddp --model_type compression_gan --regime low --batch_size 8 --n_steps 6 --log_interval 2 \
  --save gpurun_out/c5_ddp > gpurun_out/c5_train_ddp.log 2>&1; echo "train_ddp rc=$?" >> $S
ls -la gpurun_out/c5_ddp/checkpoints >> $S 2>&1
rm -rf gpurun_out/c5_ddp
cat $S
tail -5 gpurun_out/c5_train_ddp.log
python - <<'PY'
import json
for m in (1, 0):
    try:
        d = json.load(open(f"gpurun_out/c5_bench_2gpu_overlap{m}.json"))
        t, g = d["train_step"], d.get("gan_train_iteration") or {}
        print("overlap", m, "fwd img/s", round(d["value"]), "train ms", round(t["ms_per_step"], 2), t["gradient_allreduce"][:90],
              "| gan G ms", g.get("ms_per_generator_iteration"), "D ms", g.get("ms_per_discriminator_iteration"))
    except Exception as e:
        print("overlap", m, "unreadable:", e)
PY
