#!/bin/bash
# Round 2, GPU call 16: tap-grouped weight gradient of the narrow big-map layers (HFC_WGRAD_TAPG, default on): gradient
# parity (layer level, networks, GAN), training-step A/B, per-launch list of the weight-gradient kernel.
mkdir -p gpurun_out
S=gpurun_out/c16_status.txt
: > $S
T="tests/test_gpu_grad.py tests/test_gpu_gan.py tests/test_gpu_zzinstancenorm.py"
timeout 600 python -m pytest $T -m gpu -q -rfEs > gpurun_out/c16_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 500 python -m pytest tests/test_gpu_train.py -m gpu -q -rfEs -x > gpurun_out/c16_tests_train.log 2>&1; echo "train tests rc=$?" >> $S
B="--steps 24 --warmup 5 --no-gan --no-cpu-baseline --no-compress --no-eager"
timeout 200 python bench.py $B > gpurun_out/c16_bench_tapg1.json 2> gpurun_out/c16_bench_tapg1.err; echo "bench tapg=1 rc=$?" >> $S
HFC_WGRAD_TAPG=0 timeout 200 python bench.py $B > gpurun_out/c16_bench_tapg0.json 2> gpurun_out/c16_bench_tapg0.err; echo "bench tapg=0 rc=$?" >> $S
timeout 200 python tools/train_profile.py --out gpurun_out/c16_train_profile.txt --detail wgrad_igemm_kernel > /dev/null 2>&1; echo "train_profile rc=$?" >> $S
cat $S
tail -4 gpurun_out/c16_tests.log
tail -4 gpurun_out/c16_tests_train.log
python - <<'PY'
import json
for m in ("tapg1", "tapg0"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/c16_bench_{m}.json") if l.startswith("{")][-1])
        t = d["train_step"]
        print(f"{m}: fwd {d['ms_per_step']:.3f} ms train {t['ms_per_step']:.2f} ms bwd {t['phases']['backward_ms']:.2f}")
    except Exception as e:
        print(m, "unreadable", e)
PY
grep "wgrad_igemm_kernel" gpurun_out/c16_train_profile.txt | head -4 | cut -c1-120
grep -A60 "per-launch" gpurun_out/c16_train_profile.txt | awk '{print $1}' | tr '\n' ' ' | cut -c1-400
