#!/bin/bash
# Round 2, final GPU call: the whole tree as committed -- bench (all legs), ncu --set full of the fused residual conv (the
# dominant kernel of the step), smoke(), then the full GPU suite.
mkdir -p gpurun_out
S=gpurun_out/c17_status.txt
: > $S
timeout 420 python bench.py --steps 20 --warmup 5 > gpurun_out/c17_bench_full.json 2> gpurun_out/c17_bench_full.err; echo "bench full rc=$?" >> $S
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 30 -c 4 -f -o gpurun_out/r02_resconv_widenorm \
    python bench.py --profile --steps 1 --warmup 1 > gpurun_out/c17_ncu_res.log 2>&1; echo "ncu resconv rc=$?" >> $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c17_smoke.txt 2>&1; echo "smoke rc=$?" >> $S
timeout 120 python tools/layer_times.py > gpurun_out/c17_layer_times.txt 2> gpurun_out/c17_layer_times.err; echo "layer_times rc=$?" >> $S
timeout 1000 python -m pytest tests -m gpu -q -rfEs > gpurun_out/c17_tests.log 2>&1; echo "tests rc=$?" >> $S
cat $S
tail -3 gpurun_out/c17_smoke.txt
tail -6 gpurun_out/c17_tests.log
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c17_bench_full.json") if l.startswith("{")][-1])
    r = d["roofline"]
    print("fwd", round(d["ms_per_step"], 3), "ms", round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "| roofline frac", round(r["frac"], 3),
          "ms", round(r["ms_per_launch"], 4), "conv alone", round(r["conv_alone"]["frac"], 3), "step_tensor_frac", round(r["step_tensor_frac"], 3))
    print("train", d.get("train_step_ms"), "gan", d.get("gan_generator_ms"), d.get("gan_discriminator_ms"), "eager", json.dumps(d.get("eager_cudnn"))[:300])
    print("c5", json.dumps(d.get("c5_inference"))[:200])
    print("hbm", json.dumps(d.get("roofline_hbm"))[:300])
    print("cpu", json.dumps(d.get("cpu_baseline"))[:200], "clocks", d.get("clocks"))
except Exception as e:
    print("bench unreadable", e)
PY
