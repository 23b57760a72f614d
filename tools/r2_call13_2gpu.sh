#!/bin/bash
# Round 2, 2-GPU: where the in-backward reducer's time goes (phases of the training step) under a few NCCL settings.
mkdir -p gpurun_out
S=gpurun_out/c13_status.txt
: > $S
COMMON="--steps 24 --warmup 3 --no-gan --no-cpu-baseline --no-compress --no-eager"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 $COMMON > gpurun_out/c13_$name.json 2> gpurun_out/c13_$name.err
  echo "$name rc=$?" >> $S
}
timeout 300 python bench.py --gpus 1 $COMMON > gpurun_out/c13_n1.json 2> gpurun_out/c13_n1.err; echo "n1 rc=$?" >> $S
run overlap_default HFC_OVERLAP_ALLREDUCE=1
run overlap_bucket128 HFC_OVERLAP_ALLREDUCE=1 HFC_REDUCER_BUCKET_MB=128
run overlap_ctas8_hi HFC_OVERLAP_ALLREDUCE=1 HFC_REDUCER_MAX_CTAS=8 HFC_REDUCER_HIGH_PRIORITY=1
run overlap_ctas4_hi_b64 HFC_OVERLAP_ALLREDUCE=1 HFC_REDUCER_MAX_CTAS=4 HFC_REDUCER_HIGH_PRIORITY=1 HFC_REDUCER_BUCKET_MB=64
run plain HFC_OVERLAP_ALLREDUCE=0
cat $S
python - <<'PY'
import json
for f in ("n1", "overlap_default", "overlap_bucket128", "overlap_ctas8_hi", "overlap_ctas4_hi_b64", "plain"):
    try:
        lines = [l for l in open(f"gpurun_out/c13_{f}.json") if l.startswith("{")]
        d = json.loads(lines[-1]); t = d["train_step"]; ph = t.get("phases") or {}
        print(f"{f:22s} train {t['ms_per_step']:.2f} ms | fwd {ph.get('forward_and_losses_ms', 0):.2f} bwd {ph.get('backward_ms', 0):.2f} "
              f"reduce {ph.get('gradient_allreduce_after_backward_ms', 0):.2f} adam {ph.get('adam_ms', 0):.2f} host {ph.get('host_enqueue_ms', 0):.2f} | e2e {d['e2e']['value']:.0f} fwd {d['value']:.0f} | {t['gradient_allreduce'][:60]}")
    except Exception as e:
        print(f, "unreadable:", e)
PY
