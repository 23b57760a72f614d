#!/bin/bash
# Round 2, GPU call 9: fast per-tile decode (multiply-shift instead of 7 integer divisions per tile in the producer warp).
mkdir -p gpurun_out
S=gpurun_out/c9_status.txt
: > $S
timeout 60 python tools/profile_thin_layers.py > gpurun_out/c9_thin_layers.txt 2>&1; echo "thin layers rc=$?" >> $S
timeout 120 python tools/layer_times.py > gpurun_out/c9_layer_times.txt 2> gpurun_out/c9_layer_times.err; echo "layer_times rc=$?" >> $S
B="--steps 20 --warmup 5 --no-train --no-cpu-baseline --no-compress --no-eager"
timeout 150 python bench.py $B > gpurun_out/c9_bench_fwd.json 2> gpurun_out/c9_bench_fwd.err; echo "bench fwd rc=$?" >> $S
T="tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_conv_modes.py tests/test_gpu_grad.py tests/test_gpu_gan.py tests/test_gpu_zzwidenorm.py"
timeout 900 python -m pytest $T -m gpu -q -rfEs > gpurun_out/c9_tests.log 2>&1; echo "tests rc=$?" >> $S
HFC_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 12 -f -o gpurun_out/r02_bigmap_v5 \
    python tools/profile_thin_layers.py > gpurun_out/c9_ncu_thin.log 2>&1; echo "ncu thin rc=$?" >> $S
cat $S
cat gpurun_out/c9_thin_layers.txt
tail -3 gpurun_out/c9_tests.log
python -c "
import json; d=json.load(open('gpurun_out/c9_bench_fwd.json')); print('fwd', d['value'], d['ms_per_step'], d['roofline']['step_tensor_frac'])"
