#!/bin/bash
# Round 2, last GPU call: the committed tree after the optimizer refactor (Adam.step_subset) -- Adam tests, smoke, N=1 bench.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "adam" > gpurun_out/c20_tests.log 2>&1; echo "adam tests rc=$?"
tail -2 gpurun_out/c20_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c20_smoke.txt 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/c20_smoke.txt
timeout 150 python bench.py --steps 12 --warmup 3 --no-gan --no-compress --no-cpu-baseline --no-eager > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c20_bench.json") if l.startswith("{")][-1])
print("fwd", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "roofline", round(d["roofline"]["frac"], 3), "traffic", d["roofline"]["traffic"], "train", round(d["train_step"]["ms_per_step"], 2), d["train_step"]["gradient_allreduce"])
PY
