#!/bin/bash
# Round 2, GPU call 11: pipeline race fix (the formerly flaky test, 6 fresh processes), packed-fp32 thin epilogue (parity +
# timing), forward bench.
mkdir -p gpurun_out
S=gpurun_out/c11_status.txt
: > $S
pass=0; fail=0
for i in 1 2 3 4 5 6; do
  if timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "pipelined_forward" > /tmp/fp.log 2>&1; then pass=$((pass+1)); else fail=$((fail+1)); fi
done
echo "pipelined test x6: pass $pass fail $fail" >> $S
timeout 100 python tools/flaky_pipe_diag.py 2>&1 | tail -2 >> $S
timeout 60 python tools/profile_thin_layers.py > gpurun_out/c11_thin_layers.txt 2>&1; echo "thin layers rc=$?" >> $S
B="--steps 20 --warmup 5 --no-train --no-cpu-baseline --no-compress --no-eager"
timeout 150 python bench.py $B > gpurun_out/c11_bench_fwd.json 2> gpurun_out/c11_bench_fwd.err; echo "bench fwd rc=$?" >> $S
T="tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_conv_modes.py tests/test_gpu_gan.py"
timeout 900 python -m pytest $T -m gpu -q -rfEs > gpurun_out/c11_tests.log 2>&1; echo "tests rc=$?" >> $S
cat $S
cat gpurun_out/c11_thin_layers.txt
tail -3 gpurun_out/c11_tests.log
python -c "
import json; d=json.load(open('gpurun_out/c11_bench_fwd.json')); print('fwd', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['step_tensor_frac'])"
