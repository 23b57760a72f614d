#!/bin/bash
# Round 2, GPU call 14: InstanceNorm variant on the real kernels; sub-pixel phases of the transposed convs forked onto
# helper streams (HFC_PHASE_STREAMS=1, default) vs serial (=0): parity subset, layer times, forward / train bench; a
# per-launch list of the weight-packing / fold / permute kernels of one training step.
mkdir -p gpurun_out
S=gpurun_out/c14_status.txt
: > $S
timeout 300 python -m pytest tests/test_gpu_zzinstancenorm.py -m gpu -q -rfEs > gpurun_out/c14_tests_instancenorm.log 2>&1; echo "instancenorm tests rc=$?" >> $S
T="tests/test_gpu_conv_modes.py tests/test_gpu_parity.py tests/test_gpu_ops.py tests/test_gpu_grad.py tests/test_gpu_train.py"
timeout 900 python -m pytest $T -m gpu -q -rfEs -x > gpurun_out/c14_tests_phase_streams.log 2>&1; echo "tests (phase streams on) rc=$?" >> $S
for m in 1 0; do
  HFC_PHASE_STREAMS=$m timeout 120 python tools/layer_times.py > gpurun_out/c14_layer_times_ps$m.txt 2> gpurun_out/c14_layer_times_ps$m.err; echo "layer_times ps=$m rc=$?" >> $S
  HFC_PHASE_STREAMS=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-gan --no-cpu-baseline --no-compress --no-eager \
      > gpurun_out/c14_bench_ps$m.json 2> gpurun_out/c14_bench_ps$m.err; echo "bench ps=$m rc=$?" >> $S
done
timeout 200 python tools/train_profile.py --out gpurun_out/c14_train_profile.txt --detail pack_rows,pad_fold,permute_wgrad > /dev/null 2>&1; echo "train_profile rc=$?" >> $S
cat $S
tail -4 gpurun_out/c14_tests_instancenorm.log
tail -4 gpurun_out/c14_tests_phase_streams.log
python - <<'PY'
import json
for m in (1, 0):
    try:
        d = json.loads([l for l in open(f"gpurun_out/c14_bench_ps{m}.json") if l.startswith("{")][-1])
        print(f"phase streams={m}: fwd {d['ms_per_step']:.3f} ms ({d['value']:.0f} img/s) e2e {d['e2e']['value']:.0f} train {d['train_step']['ms_per_step']:.2f} ms tensor_frac {d['roofline'].get('step_tensor_frac')}")
    except Exception as e:
        print(m, "unreadable", e)
    try:
        for l in open(f"gpurun_out/c14_layer_times_ps{m}.txt"):
            if l.startswith(("H.", "G.up", "Model", "Generator total", "Encoder total")):
                print("   ", l.rstrip()[:100])
    except Exception as e:
        print(m, "no layer times", e)
PY
