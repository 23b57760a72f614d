#!/bin/bash
# Round 2, GPU call 1: run everything round 1 left un-run + first backward-precision table (item 2).
#   gpurun --timeout 1500 -- 'bash tools/r2_call1.sh'
mkdir -p gpurun_out
S=gpurun_out/c1_status.txt
: > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
# 1. the whole GPU suite, gates open, NOT -x: every failure is information
HFC_RUN_UNVERIFIED=1 timeout 600 python -m pytest tests -m gpu -q -rfEs > gpurun_out/c1_tests.log 2>&1; echo "tests rc=$?" >> $S
# 2. backward precision, current format (bf16 operands)
timeout 240 python tools/grad_precision.py > gpurun_out/c1_grad_precision.txt 2>&1; echo "grad_precision rc=$?" >> $S
# 3. thin epilogue: parity through the op / model tests with the switch set, then layer times on / off
HFC_THIN_EPILOGUE=1 timeout 240 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_conv_modes.py -q -rfE \
    > gpurun_out/c1_tests_thin.log 2>&1; echo "tests thin rc=$?" >> $S
timeout 60 python tools/profile_thin_layers.py > gpurun_out/c1_thin_layers_off.txt 2>&1
HFC_THIN_EPILOGUE=1 timeout 60 python tools/profile_thin_layers.py > gpurun_out/c1_thin_layers_on.txt 2>&1; echo "thin layers rc=$?" >> $S
# 4. forward bench: default / fused 960-wide norm / thin / both
B="--steps 20 --warmup 5 --no-train --no-cpu-baseline --no-compress"
timeout 150 python bench.py $B > gpurun_out/c1_bench_fwd.json 2> gpurun_out/c1_bench_fwd.err; echo "bench fwd rc=$?" >> $S
HFC_FUSE_RESNORM=1 timeout 150 python bench.py $B > gpurun_out/c1_bench_fwd_resnorm.json 2> gpurun_out/c1_bench_fwd_resnorm.err; echo "bench resnorm rc=$?" >> $S
HFC_THIN_EPILOGUE=1 timeout 150 python bench.py $B > gpurun_out/c1_bench_fwd_thin.json 2> gpurun_out/c1_bench_fwd_thin.err; echo "bench thin rc=$?" >> $S
HFC_THIN_EPILOGUE=1 HFC_FUSE_RESNORM=1 timeout 150 python bench.py $B > gpurun_out/c1_bench_fwd_both.json 2> gpurun_out/c1_bench_fwd_both.err; echo "bench both rc=$?" >> $S
# 5. full bench (train step incl. native LPIPS trunk timing)
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench_full.json 2> gpurun_out/c1_bench_full.err; echo "bench full rc=$?" >> $S
cat $S
tail -5 gpurun_out/c1_tests.log
cat gpurun_out/c1_grad_precision.txt | tail -12
cat gpurun_out/c1_thin_layers_off.txt gpurun_out/c1_thin_layers_on.txt
