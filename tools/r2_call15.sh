#!/bin/bash
# Round 2, GPU call 15: vectorised weight packing / weight-gradient permute, multi-group ChannelNorm rows kernel
# (parity subset + training-step A/B), InstanceNorm network test against the operand-matched oracle.
mkdir -p gpurun_out
S=gpurun_out/c15_status.txt
: > $S
timeout 300 python -m pytest tests/test_gpu_zzinstancenorm.py -m gpu -q -rfEs -s > gpurun_out/c15_tests_instancenorm.log 2>&1; echo "instancenorm tests rc=$?" >> $S
T="tests/test_gpu_ops.py tests/test_gpu_conv_modes.py tests/test_gpu_grad.py tests/test_gpu_gan.py"
timeout 600 python -m pytest $T -m gpu -q -rfEs -x > gpurun_out/c15_tests.log 2>&1; echo "tests rc=$?" >> $S
B="--steps 24 --warmup 5 --no-gan --no-cpu-baseline --no-compress --no-eager"
timeout 200 python bench.py $B > gpurun_out/c15_bench_new.json 2> gpurun_out/c15_bench_new.err; echo "bench new rc=$?" >> $S
HFC_PACK_VEC=0 HFC_PERMUTE_VEC=0 HFC_CN_ITER=0 timeout 200 python bench.py $B > gpurun_out/c15_bench_old.json 2> gpurun_out/c15_bench_old.err; echo "bench old rc=$?" >> $S
timeout 200 python tools/train_profile.py --out gpurun_out/c15_train_profile.txt > /dev/null 2>&1; echo "train_profile rc=$?" >> $S
cat $S
tail -5 gpurun_out/c15_tests_instancenorm.log
grep "worst parameter gradient" gpurun_out/c15_tests_instancenorm.log
tail -4 gpurun_out/c15_tests.log
python - <<'PY'
import json
for m in ("new", "old"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/c15_bench_{m}.json") if l.startswith("{")][-1])
        t = d["train_step"]
        print(f"{m}: fwd {d['ms_per_step']:.3f} ms e2e {d['e2e']['value']:.0f} train {t['ms_per_step']:.2f} ms phases {t.get('phases')}")
    except Exception as e:
        print(m, "unreadable", e)
PY
head -24 gpurun_out/c15_train_profile.txt | cut -c1-120
